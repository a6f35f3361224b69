// K6: mix-bus exchange between the GPUs of one box over peer memory (NVLink / NVSwitch), fused into the kernel
// that reduces the per-warp mix partials -- the path's only inter-GPU step (16 KiB per block).
#pragma once

#include "common.cuh"

namespace mxb {

constexpr int kExchMaxWorld = 16;
constexpr int kExchFlagBytes = 128;          // one flag per slot, alone in its line

struct ExchDev {                              // passed to the kernel by value
    double* local_payload;                    // my slot of this call
    unsigned long long* local_flag;
    const double* peer_payload[kExchMaxWorld];          // same slot in every rank's buffer (index = rank; own entry = local)
    const unsigned long long* peer_flag[kExchMaxWorld];
    int rank, world;
    unsigned long long seq1;                  // value the flags reach when this call's payload is published
    unsigned int* ticket;                     // last-CTA election
};

}  // namespace mxb

struct mxb_exchange {
    mxb_ctx* ctx;
    int rank, world, max_doubles;
    size_t slot_bytes;                        // flag line + payload, 128-byte multiple
    unsigned char* local;                     // cudaMalloc'ed: 2 slots
    unsigned char* peers[mxb::kExchMaxWorld]; // cudaIpcOpenMemHandle'd (own entry = local)
    bool connected;
    unsigned long long seq;                   // calls so far
    unsigned int* ticket;
};

namespace mxb {
// fills the device descriptor for the next call and advances the sequence
ExchDev exchange_next(mxb_exchange* ex);
}
