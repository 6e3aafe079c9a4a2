#pragma once
#include "common.cuh"

struct LogmelTables {
    int n_mels;
    float* dft;       // [400][416] windowed DFT matrix (cos | -sin), n-major
    float* fb_w;      // banded slaney filterbank weights
    int* fb_meta;     // per mel bin: (first fft bin, band length, offset into fb_w)
};

int logmel_tables_create(int n_mels, LogmelTables* t);
void logmel_tables_destroy(LogmelTables* t);
// audio f32[B,480000] -> mel f32[B,n_mels,3000]; clipmax_scratch: device f32[B]
int launch_logmel(const LogmelTables& t, const float* audio, int B, float* mel, float* clipmax_scratch,
                  cudaStream_t st);
