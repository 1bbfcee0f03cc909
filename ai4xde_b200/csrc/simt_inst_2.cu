// simt_inst_2.cu -- explicit instantiations of the FP32 engine (split for parallel builds).
#include "simt_kernels.cuh"
namespace ax {
template int simt_forward<P_3>(const FwdArgs &);
template int simt_backward<P_3>(const BwdArgs &);
template int simt_forward<P_3_D>(const FwdArgs &);
template int simt_backward<P_3_D>(const BwdArgs &);
template int simt_forward<P_3_F>(const FwdArgs &);
template int simt_backward<P_3_F>(const BwdArgs &);
}  // namespace ax
