/*
 * ec_pmon.h -- energy / power sampling of the GPUs the driver runs on, successor of the reference's EC_PMON hook
 * (src/common/module/ec_pmon_mod.F90:14-57 reads the Cray /sys/cray/pm_counters/{energy,power} files; sampled inside
 * the block loop and summarised after it, src/cloudsc_fortran/cloudsc_driver_mod.F90:113-118,170-178,190-200).
 * Here the counters are NVML's (nvmlDeviceGetTotalEnergyConsumption [mJ], nvmlDeviceGetPowerUsage [mW]); the library
 * is dlopen'ed, so the driver starts without it.  Enabled, as upstream, by EC_PMON=1 in the environment.
 */
#ifndef CLOUDSC_EC_PMON_H
#define CLOUDSC_EC_PMON_H

/* 1 if EC_PMON=1 and NVML could be loaded and initialised, else 0 (then every call below is a no-op returning 0) */
int ec_pmon_init(void);
/* energy counter of GPU `index` in joules since driver load, and its present power draw in watts; 0 on success */
int ec_pmon(int index, double *energy_j, double *power_w);
void ec_pmon_shutdown(void);
#endif
