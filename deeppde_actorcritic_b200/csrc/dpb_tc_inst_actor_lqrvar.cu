// actor_tc_kernel<24, EQ_LQRVAR, 0> (see dpb_tc_inst.cuh)
#define DPB_INST_NAME actor_lqrvar
#define DPB_INST_KERNEL actor_tc_kernel
#define DPB_INST_DP 24
#define DPB_INST_EQN EQ_LQRVAR
#define DPB_INST_MV 0
#include "dpb_tc_inst.cuh"
