// Tile geometry shared by the planner (work-unit sizing) and the pair kernels.
#pragma once
#define PAIR_NT 256      // threads per CTA
#define PAIR_TM 1024     // targets per tile
#define PAIR_SMAX 2048   // staged source events per tile (16 KB)
