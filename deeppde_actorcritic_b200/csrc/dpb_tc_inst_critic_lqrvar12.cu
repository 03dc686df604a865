// critic_tc_kernel<12, EQ_LQRVAR, 0> (see dpb_tc_inst.cuh)
#define DPB_INST_NAME critic_lqrvar12
#define DPB_INST_KERNEL critic_tc_kernel
#define DPB_INST_DP 12
#define DPB_INST_EQN EQ_LQRVAR
#define DPB_INST_MV 0
#include "dpb_tc_inst.cuh"
