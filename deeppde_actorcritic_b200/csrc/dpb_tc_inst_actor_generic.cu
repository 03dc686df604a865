// actor_tc_kernel<0, -1, 0> (see dpb_tc_inst.cuh)
// helper groups of the actor kernel: 0 = the owners do the helpers' work themselves (dpb_tc_nets.cuh)
#ifndef DPB_TC_NGRP
#define DPB_TC_NGRP 0
#endif
#define DPB_INST_NAME actor_generic
#define DPB_INST_KERNEL actor_tc_kernel
#define DPB_INST_DP 0
#define DPB_INST_EQN -1
#define DPB_INST_MV 0
#include "dpb_tc_inst.cuh"
