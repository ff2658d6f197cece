#define FDR_GROUP_LOGNS X(11) X(12)
#include "passes_group.inc"
