// Configuration unit "oxford12" of libwsi_b200.so: the kernels of wsi_kernels.cuh instantiated for one
// pollutant configuration (see wsi_host.h).
#include "wsi_kernels.cuh"

WSI_DEFINE_CONFIG_UNIT(oxford12, (runtime_cfg_as<CfgStaticMask<12, kOxfordMask>>(d)))
