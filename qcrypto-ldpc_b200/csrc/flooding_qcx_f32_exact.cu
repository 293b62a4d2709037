// clustered flooding kernel (flooding_qcx_impl.cuh), instantiation: f32_exact
#include "flooding_qcx_impl.cuh"

QL_QCX_DEFINE(f32_exact, float, float, float, kSpaExact)
