#include "zb_internal.h"
namespace zb { int deflate_init(zb200_ctx *) { return ZB200_OK; } }
