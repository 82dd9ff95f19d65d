/*
 * TEST-ONLY helper: the product's batch engine (rhfq_engine.inl) compiled with
 * RHFQ_EMU, i.e. every "kernel" runs as a host loop.  It exists to debug the
 * level-synchronous pipeline logic in the build container, which has no GPU.
 * It is never built into or loaded by the package (symbols are emu_rhfq_*).
 */
#define RHFQ_EMU 1
#include "../../reheatfunq_b200/csrc/rhfq_engine.inl"
#include "../../reheatfunq_b200/csrc/rhfq_resilience.inl"
#include "../../reheatfunq_b200/csrc/rhfq_gcp.inl"
#include "../../include/reheatfunq_b200.h"
#define RHFQ_API(name) emu_rhfq_##name
#include "../../reheatfunq_b200/csrc/rhfq_capi.inl"
