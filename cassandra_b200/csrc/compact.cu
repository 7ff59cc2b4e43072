// compact.cu — b200c_compact (placeholder until the merge pipeline lands in this file), poll and cancel.
#include "engine.cuh"
using namespace b200c;
extern "C" {
int b200c_compact(b200c_ctx* c, const b200c_manifest* m, b200c_result* r, int flags) {
    if (!c || !m || !r) return B200C_EINVAL;
    c->err = "b200c_compact: not implemented yet"; return B200C_EUNSUPPORTED;
}
int b200c_poll(b200c_ctx* c, b200c_progress* p) {
    if (!c || !p) return B200C_EINVAL;
    p->bytes_scanned = c->prog_scanned.load(); p->bytes_total = c->prog_total.load(); p->stage = c->prog_stage.load(); p->_pad = 0;
    return B200C_OK;
}
void b200c_cancel(b200c_ctx* c) { if (c) c->cancel.store(1); }
}
