// simt_inst_1.cu -- explicit instantiations of the FP32 engine (split for parallel builds).
#include "simt_kernels.cuh"
namespace ax {
template int simt_forward<P_2_A>(const FwdArgs &);
template int simt_backward<P_2_A>(const BwdArgs &);
template int simt_forward<P_2_D>(const FwdArgs &);
template int simt_backward<P_2_D>(const BwdArgs &);
template int simt_forward<P_2_F>(const FwdArgs &);
template int simt_backward<P_2_F>(const BwdArgs &);
template int simt_forward<P_2_3>(const FwdArgs &);
template int simt_backward<P_2_3>(const BwdArgs &);
}  // namespace ax
