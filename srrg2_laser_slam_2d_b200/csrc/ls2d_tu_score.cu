// ls2d_tu_score.cu -- the scoring pass (ls2d_score_batch): placeholder translation unit, see ls2d_tu_icp.cu
#include "ls2d_internal.h"
