// simt_inst_0.cu -- explicit instantiations of the FP32 engine (split for parallel builds).
#include "simt_kernels.cuh"
namespace ax {
template int simt_forward<P_V>(const FwdArgs &);
template int simt_backward<P_V>(const BwdArgs &);
template int simt_forward<P_1>(const FwdArgs &);
template int simt_backward<P_1>(const BwdArgs &);
template int simt_forward<P_1_D>(const FwdArgs &);
template int simt_backward<P_1_D>(const BwdArgs &);
template int simt_forward<P_2>(const FwdArgs &);
template int simt_backward<P_2>(const BwdArgs &);
template int simt_forward<P_2_T>(const FwdArgs &);
template int simt_backward<P_2_T>(const BwdArgs &);
template int simt_forward<P_3_T>(const FwdArgs &);
template int simt_backward<P_3_T>(const BwdArgs &);
template int simt_forward<P_1_3>(const FwdArgs &);
template int simt_backward<P_1_3>(const BwdArgs &);
template int simt_forward<P_1_4>(const FwdArgs &);
template int simt_backward<P_1_4>(const BwdArgs &);
}  // namespace ax
