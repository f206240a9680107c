#pragma once
namespace ipcr {
static bool jit_try_launch(ipcr_scratch *, const GenomeData &, int, uint64_t) { return false; }
}
