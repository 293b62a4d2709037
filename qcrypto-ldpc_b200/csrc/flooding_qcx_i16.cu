// clustered flooding kernel (flooding_qcx_impl.cuh), instantiation: i16
#include "flooding_qcx_impl.cuh"

QL_QCX_DEFINE(i16, int16_t, int, int16_t, kMinSum)
