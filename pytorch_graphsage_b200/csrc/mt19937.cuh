// mt19937.cuh -- the device MT19937 stream object (see mt19937.cu for the layout)
#pragma once
#include "common.cuh"

struct gsage_rng {
    uint32_t* ring = nullptr;      // device, cap untempered words
    int64_t cap = 0;               // power of two
    int64_t* cursor = nullptr;     // device int64[2], double-buffered
    int* err_flag = nullptr;       // device, sticky
    int* tile_count = nullptr;     // device scratch
    int64_t* tile_off = nullptr;   // device scratch
    int tiles_cap = 0;
    int parity = 0;                // host: which cursor slot is current
    int64_t gen_end = 0;           // host: generation of words [.., gen_end) is queued (on `side`)
    int64_t gen_visible = 0;       // host: words the caller's stream has already been ordered after
    int64_t max_window = 0;        // host: largest look-ahead window requested since the last seed (prefetch size)
    cudaStream_t side = nullptr;   // refills run here, overlapping the caller's compute
    cudaStream_t side_now = nullptr;
    bool overlap = true;           // GSAGE_RNG_OVERLAP=0 keeps refills on the caller's stream
    cudaEvent_t ev_main = nullptr, ev_refill = nullptr;
    int64_t cursor_lb = 0, cursor_ub = 0;   // host bounds on the device cursor
    int64_t origin = 0;            // cursor value at the last seed / set_state
    int64_t prefetch_blocks = 64;  // sequential refill: generate at least this many 624-word blocks
    // lane-parallel refill (jump-ahead): lanes x lane_blocks blocks per refill
    int lanes = 32, lane_blocks = 256;
    int64_t lane_threshold = 128;  // refills needing fewer blocks than this stay on the sequential kernel
    bool lanes_ready = false;
    uint32_t* polys = nullptr;     // device, (lanes-1) x 624 words: t^(l*lane_blocks*624) mod phi
    uint32_t* partial = nullptr;   // device, (lanes-1) x 8 x 624 words
    // consumers may alternate between streams (the engine's sample-ahead stream and the caller's): the first draw on a
    // new stream is ordered after everything queued on the previous one, so the stream of draws stays one sequence
    cudaStream_t last_stream = nullptr; bool have_last = false;
    cudaEvent_t ev_switch = nullptr;
    // asynchronous cursor feedback: after a draw the true cursor is copied to pinned host memory (8 bytes, stream ordered);
    // once the copy has landed the host bounds are re-tightened from it WITHOUT synchronising anything.  Accepted draws are
    // only a lower bound on words consumed (masked rejection), so without this the bounds drift apart by ~40 % of every
    // draw and the ring-room check forces a blocking read-back every few batches.
    static constexpr int kFeedback = 4;
    struct Feedback { int64_t* host = nullptr; cudaEvent_t ev = nullptr; int64_t acc_at = 0, win_at = 0; bool pending = false; } fb[kFeedback];
    int64_t acc_total = 0, win_total = 0;     // monotone: words certainly / at most consumed by all draws queued so far
};

namespace gsage {
// bounded draws into a device buffer (gsage_rng_randint without the argument checks)
int rng_randint_internal(gsage_rng* r, uint32_t hi, int64_t count, uint32_t* out, cudaStream_t s);
// order stream `s` after the stream that drew from `r` last (no-op when it is the same stream)
int rng_enter(gsage_rng* r, cudaStream_t s);
}
