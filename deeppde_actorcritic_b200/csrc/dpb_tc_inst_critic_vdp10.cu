// critic_tc_kernel<20, EQ_VDP, 10> (see dpb_tc_inst.cuh)
#define DPB_INST_NAME critic_vdp10
#define DPB_INST_KERNEL critic_tc_kernel
#define DPB_INST_DP 20
#define DPB_INST_EQN EQ_VDP
#define DPB_INST_MV 10
#include "dpb_tc_inst.cuh"
