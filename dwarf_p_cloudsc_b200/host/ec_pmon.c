/* ec_pmon.c -- see ec_pmon.h */
#include "ec_pmon.h"

#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>

typedef int (*nvml_init_t)(void);
typedef int (*nvml_shutdown_t)(void);
typedef int (*nvml_handle_t)(unsigned, void **);
typedef int (*nvml_energy_t)(void *, unsigned long long *);
typedef int (*nvml_power_t)(void *, unsigned *);

static void *g_lib = NULL;
static nvml_shutdown_t g_shutdown = NULL;
static nvml_handle_t g_handle = NULL;
static nvml_energy_t g_energy = NULL;
static nvml_power_t g_power = NULL;

int ec_pmon_init(void) {
  const char *e = getenv("EC_PMON");
  if (!e || strcmp(e, "1") != 0) return 0;   /* as upstream: only the value '1' switches it on */
  if (g_lib) return 1;
  void *lib = dlopen("libnvidia-ml.so.1", RTLD_NOW | RTLD_LOCAL);
  if (!lib) return 0;
  nvml_init_t init = (nvml_init_t)dlsym(lib, "nvmlInit_v2");
  g_shutdown = (nvml_shutdown_t)dlsym(lib, "nvmlShutdown");
  g_handle = (nvml_handle_t)dlsym(lib, "nvmlDeviceGetHandleByIndex_v2");
  g_energy = (nvml_energy_t)dlsym(lib, "nvmlDeviceGetTotalEnergyConsumption");
  g_power = (nvml_power_t)dlsym(lib, "nvmlDeviceGetPowerUsage");
  if (!init || !g_handle || !g_energy || !g_power || init() != 0) { dlclose(lib); return 0; }
  g_lib = lib;
  return 1;
}

int ec_pmon(int index, double *energy_j, double *power_w) {
  if (energy_j) *energy_j = 0.0;
  if (power_w) *power_w = 0.0;
  if (!g_lib) return 0;
  void *dev = NULL;
  if (g_handle((unsigned)index, &dev) != 0) return -1;
  unsigned long long mj = 0;
  unsigned mw = 0;
  const int re = g_energy(dev, &mj), rp = g_power(dev, &mw);
  if (energy_j && re == 0) *energy_j = 1.0e-3 * (double)mj;
  if (power_w && rp == 0) *power_w = 1.0e-3 * (double)mw;
  return (re == 0 || rp == 0) ? 0 : -1;
}

void ec_pmon_shutdown(void) {
  if (!g_lib) return;
  if (g_shutdown) g_shutdown();
  dlclose(g_lib);
  g_lib = NULL;
}
