#define FDR_GROUP_LOGNS X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8)
#include "passes_group.inc"
