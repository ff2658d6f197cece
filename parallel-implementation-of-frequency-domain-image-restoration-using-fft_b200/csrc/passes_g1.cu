#define FDR_GROUP_LOGNS X(9) X(10)
#include "passes_group.inc"
