#define RFB_RM_DOUBLE 1
#include "regmix_launch.cu"
