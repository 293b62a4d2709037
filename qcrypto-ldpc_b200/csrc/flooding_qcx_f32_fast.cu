// clustered flooding kernel (flooding_qcx_impl.cuh), instantiation: f32_fast
#include "flooding_qcx_impl.cuh"

QL_QCX_DEFINE(f32_fast, float, float, float, kSpaFast)
