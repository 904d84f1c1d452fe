#include "zb_internal.h"
namespace zb { int inflate_init(zb200_ctx *) { return ZB200_OK; } }
