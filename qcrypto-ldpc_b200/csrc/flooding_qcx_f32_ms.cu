// clustered flooding kernel (flooding_qcx_impl.cuh), instantiation: f32_ms
#include "flooding_qcx_impl.cuh"

QL_QCX_DEFINE(f32_ms, float, float, float, kMinSum)
