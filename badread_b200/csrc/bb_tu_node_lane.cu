// bb_tu_node_lane.cu — compiles the lane-mode node kernel bb_k_node_lane<8> (bb_tasks.cuh).
#include "bb_launch.h"

void bbl_node_lane8(int grid, cudaStream_t st, BBBatchDev B, BBQueues Q, int parity, int *cursor) {
    bb_k_node_lane<BB_NODE_LW_SMALL><<<grid, 64, 0, st>>>(B, Q, parity, cursor);
}
