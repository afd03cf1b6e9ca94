// Configuration unit "runtime15" of libwsi_b200.so: the kernels of wsi_kernels.cuh instantiated for one
// pollutant configuration (see wsi_host.h).  Models with 13 or 14 pollutants are padded up to this one.
#include "wsi_kernels.cuh"

WSI_DEFINE_CONFIG_UNIT(runtime15, runtime_cfg<15>(d))
