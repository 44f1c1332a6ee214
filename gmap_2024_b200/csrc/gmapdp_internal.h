/* What the chaining engine (gmapchain_kernels.cu) needs from the context that gmapdp_kernels.cu owns. */
#ifndef GMAPDP_INTERNAL_H
#define GMAPDP_INTERNAL_H

#include <string>
#include "../../include/gmapdp_b200.h"

struct GdpCtxView {
  int device, sm_count;
  std::string *err;
  long *launches;
  void **chain;
  void (**chain_free) (void *);
};
GdpCtxView gmapdp_ctx_view (gmapdp_ctx *ctx);

#endif
