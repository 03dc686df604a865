// dpb_tc_inst.cuh -- one translation unit per instantiation of the tensor-path kernels, so that the library builds in
// parallel (each instantiation is ~20 s of ptxas).  A unit defines DPB_INST_NAME / _KERNEL / _DP / _EQN / _MV and
// includes this file; dpb_api.cu fetches the kernel through the getter declared in DPB_TC_GETTERS.
#include "dpb_tc_kernels.cuh"
#define DPB_CAT2(a, b) a##b
#define DPB_CAT(a, b) DPB_CAT2(a, b)
namespace dpb {
namespace tc {
TcKernelFn DPB_CAT(tc_get_, DPB_INST_NAME)() { return DPB_INST_KERNEL<DPB_INST_DP, DPB_INST_EQN, DPB_INST_MV>; }
}  // namespace tc
}  // namespace dpb
