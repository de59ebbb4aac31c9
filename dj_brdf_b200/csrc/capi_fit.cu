// capi_fit.cu -- C-ABI entry points of the power-iteration fits (placeholder until kernels_fit.cu lands)
#include "djb_internal.h"
extern "C" {
djb200_status djb200_fit_tabular(const djb200_source *, int32_t, int32_t, int32_t, int32_t, djb200_tabular_fit *, void *)
{
	return DJB200_ERR_UNSUPPORTED;
}
djb200_status djb200_fit_tabular_anisotropic(const djb200_source *, int32_t, int32_t, int32_t, int32_t, int32_t,
                                             djb200_tabular_anisotropic_fit *, void *)
{
	return DJB200_ERR_UNSUPPORTED;
}
}
