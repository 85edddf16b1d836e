/*
 * dwarf_cloudsc.c -- main() of dwarf-cloudsc-b200, same command line as dwarf-cloudsc-c
 * (reference: src/cloudsc_c/dwarf_cloudsc.c):   dwarf-cloudsc-b200 [NUMOMP NGPTOT NPROMA]
 * NUMOMP selects how many GPUs are driven (<= 0: all visible); defaults 1 / 100 / 4 as upstream.
 */
#include <stdio.h>
#include <stdlib.h>

#include "cloudsc_driver.h"

int main(int argc, char *argv[]) {
  int omp_threads = 1, ngptot = 100, nproma = 4;
  if (argc == 4) {
    omp_threads = atoi(argv[1]);
    ngptot = atoi(argv[2]);
    nproma = atoi(argv[3]);
  } else if (argc != 1) {
    printf("Calling c-cloudsc with the right number of arguments will work better ;-) \n");
    return EXIT_FAILURE;
  }
  cloudsc_driver(omp_threads, ngptot, nproma);
  return cloudsc_driver_status() ? EXIT_FAILURE : 0;
}
