// critic_tc_kernel<0, -1, 0> (see dpb_tc_inst.cuh)
#define DPB_INST_NAME critic_generic
#define DPB_INST_KERNEL critic_tc_kernel
#define DPB_INST_DP 0
#define DPB_INST_EQN -1
#define DPB_INST_MV 0
#include "dpb_tc_inst.cuh"
