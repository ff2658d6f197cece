#define FDR_GROUP_LOGNS X(13) X(14)
#include "passes_group.inc"
