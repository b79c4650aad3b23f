// Shared between the two translation units of libdabmod_b200.so (dabmod_b200.cu: the modulator,
// coder_b200.cu: the channel coding ahead of it).  Not part of the C ABI.
#pragma once

#include <cuda_runtime.h>

#include <functional>
#include <stdexcept>
#include <string>

#include "../../include/dabmod_b200.h"

namespace dabmod {

// Carries a DABMOD_B200_E* code to the C boundary, where `guard` turns it into the return value.
struct ApiError : std::runtime_error {
    int code;
    ApiError(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

// What sits in front of the symbol kernels in the sliced host pipeline
//   H2D(i+1) on s_in | front + modulator kernels(i) on s_compute | D2H(i-1) on s_out:
// `upload` enqueues the host->device copy of the slice's input (TFs [t0, t0+nt) of the call) on s_in,
// `encode` enqueues whatever turns that input into BlockPartitioner blocks on the compute stream and returns
// the device pointer of the slice's first block (a plain pointer offset when the input already is blocks).
struct PipeFront {
    std::function<void(size_t t0, size_t nt, cudaStream_t s_in)> upload;
    std::function<const uint8_t *(size_t t0, size_t nt, cudaStream_t s_compute)> encode;
};

// Where the I/Q goes: a host buffer (pinned or pageable), or a file descriptor through the handle's pinned ring.
struct PipeSink {
    bool to_fd = false;
    void *host_out = nullptr;
    size_t cap = 0;
    int fd = -1;
};

// n_tf transmission frames of the handle's stream through the sliced pipeline.  Takes the handle's lock.
// Throws ApiError.  *out_bytes = bytes delivered (also on a failing sink).
void run_pipeline(dabmod_b200 *h, size_t n_tf, const PipeFront &front, const PipeSink &sink, size_t *out_bytes);

// dabmod_b200_seek with the previous TF's blocks already on the device (nullptr = stream start); enqueued on the
// handle's compute stream and synchronised.  Takes the handle's lock.
void seek_device(dabmod_b200 *h, uint64_t tf_index, const uint8_t *d_prev_bits);

// The handle's compute stream (the coder front enqueues on it) and device ordinal.
cudaStream_t compute_stream(dabmod_b200 *h);
int device_of(const dabmod_b200 *h);

} // namespace dabmod
