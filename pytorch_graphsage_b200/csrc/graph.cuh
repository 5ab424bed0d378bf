// graph.cuh -- the device adjacency object shared by graph.cu / sample.cu / engine.cu
#pragma once
#include "common.cuh"
#include <vector>

struct gsage_graph {
    int64_t n_rows = 0, n_cols = 0, nnz = 0;
    bool fast = false;         // columns redundant + no stored zeros: degree = indptr pair
    bool val64 = false;        // values stored as int64 (else int32)
    int64_t* indptr = nullptr; // device
    void* val = nullptr;       // device, int32 or int64 [nnz]
    int32_t* col = nullptr;    // device, general graphs only
    int32_t* deg = nullptr;    // device, general graphs only (non-zero count per row)
    int* err_flag = nullptr;   // device, sticky "id out of range"
    int64_t device_bytes = 0;
    std::vector<int32_t> host_deg;
};
