// trace_inst.cu -- the trace-kernel instantiations of ONE <PATHLEN, HASDET> pair (-DSMCRT_INST_PL=0|1 -DSMCRT_INST_HD=0|1):
// compiled four times, in parallel, by rsmcrt_b200/build.py (13 kernels each; one translation unit took 2.5 minutes).
#define SMCRT_TRACE_TU 1
#include "kernels.cuh"

#ifndef SMCRT_INST_PL
#error "compile with -DSMCRT_INST_PL=0|1 -DSMCRT_INST_HD=0|1"
#endif
#define SMCRT_CAT2(a, b, c, d) a##b##c##d
#define SMCRT_CAT(a, b, c, d) SMCRT_CAT2(a, b, c, d)

namespace smcrt_dev {

trace_kernel_t SMCRT_CAT(pick_kernel_pl, SMCRT_INST_PL, _hd, SMCRT_INST_HD)(int sched, int mb, bool need, bool simple, bool lean) {
    constexpr bool PL = SMCRT_INST_PL != 0, HD = SMCRT_INST_HD != 0;
    if (sched == SCHED_QUEUED) {
        if (simple && lean) return mb == 2 ? trace_queued<PL, HD, 2, true, true> : trace_queued<PL, HD, 3, true, true>;
        if (simple) return mb == 2 ? trace_queued<PL, HD, 2, true, false> : trace_queued<PL, HD, 3, true, false>;
        return mb == 2 ? trace_queued<PL, HD, 2, false, false> : trace_queued<PL, HD, 3, false, false>;
    }
    if (sched == SCHED_COMPACT) return trace_persistent<PL, HD, true, 2, false>;
    if (need) return mb == 2 ? trace_persistent<PL, HD, false, 2, true> : (mb == 4 ? trace_persistent<PL, HD, false, 4, true> : trace_persistent<PL, HD, false, 3, true>);
    return mb == 2 ? trace_persistent<PL, HD, false, 2, false> : (mb == 4 ? trace_persistent<PL, HD, false, 4, false> : trace_persistent<PL, HD, false, 3, false>);
}

}  // namespace smcrt_dev
