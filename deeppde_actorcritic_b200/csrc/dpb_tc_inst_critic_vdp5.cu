// critic_tc_kernel<12, EQ_VDP, 5> (see dpb_tc_inst.cuh)
#define DPB_INST_NAME critic_vdp5
#define DPB_INST_KERNEL critic_tc_kernel
#define DPB_INST_DP 12
#define DPB_INST_EQN EQ_VDP
#define DPB_INST_MV 5
#include "dpb_tc_inst.cuh"
