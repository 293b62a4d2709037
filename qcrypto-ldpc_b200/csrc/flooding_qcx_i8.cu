// clustered flooding kernel (flooding_qcx_impl.cuh), instantiation: i8
#include "flooding_qcx_impl.cuh"

QL_QCX_DEFINE(i8, int8_t, int16_t, int8_t, kMinSum)
