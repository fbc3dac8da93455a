// cobs/settings.hpp -- process-wide options, same names as the reference (cobs/settings.hpp:15-24)
// plus the GPU placement knobs of the B200 build.
#pragma once
#include <cstddef>

namespace cobs {

//! kept for source compatibility; the B200 path does not use host threads for the search
extern size_t gopt_threads;
//! kept for source compatibility; the index is always loaded completely -- into HBM
extern bool gopt_load_complete_index;
//! unused by the query path (FastA/FastQ caches belong to construction)
extern bool gopt_disable_cache;

//! first CUDA device to use (also read from $COBS_GPU_DEVICE)
extern int gopt_gpu_device;
//! number of GPUs an index is sharded over along the document axis (also $COBS_GPUS)
extern unsigned gopt_gpus;

} // namespace cobs
