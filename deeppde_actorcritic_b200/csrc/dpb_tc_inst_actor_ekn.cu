// actor_tc_kernel<24, EQ_EKN, 0> (see dpb_tc_inst.cuh)
// one thread per path (255 registers/thread): measured 7 % faster than two threads per path for the actor kernel
#ifndef DPB_TC_NGRP
#define DPB_TC_NGRP 1
#endif
#define DPB_INST_NAME actor_ekn
#define DPB_INST_KERNEL actor_tc_kernel
#define DPB_INST_DP 24
#define DPB_INST_EQN EQ_EKN
#define DPB_INST_MV 0
#include "dpb_tc_inst.cuh"
