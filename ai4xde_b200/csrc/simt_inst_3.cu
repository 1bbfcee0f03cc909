// simt_inst_3.cu -- explicit instantiations of the FP32 engine (split for parallel builds).
#include "simt_kernels.cuh"
namespace ax {
template int simt_forward<P_4>(const FwdArgs &);
template int simt_backward<P_4>(const BwdArgs &);
template int simt_forward<P_4_D>(const FwdArgs &);
template int simt_backward<P_4_D>(const BwdArgs &);
template int simt_forward<P_4_F>(const FwdArgs &);
template int simt_backward<P_4_F>(const BwdArgs &);
}  // namespace ax
