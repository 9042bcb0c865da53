// mcw_kernels.h -- host-callable launchers of the kernels in mcw_kernels.cu
#pragma once
#include <cuda_runtime.h>

#include "mcw_types.h"

void mcwk_init_histories(const McwParams& P, const mcw_source_site* src, int64_t n, int32_t* q_fly, cudaStream_t st);
void mcwk_sample_source(const McwParams& P, mcw_source_site* out, int64_t n, const double* d_box, int only_fissionable,
                        cudaStream_t st);
void mcwk_xs_lookup_array(const McwLibraryDev& L, int64_t n, const double* E, const int32_t* mat, double* out,
                          cudaStream_t st);
int mcwk_xs_refresh(const McwParams& P, const int32_t* queue, const unsigned long long* count, int64_t upper, bool binned,
                    const uint32_t* keys, const int32_t* hist, uint32_t* bins, uint32_t* bin_offsets, int32_t* hist_binned,
                    cudaStream_t st);
int mcwk_refresh_bins(void); // number of energy buckets of the binned refresh
void mcwk_xs_sort(int64_t n, const double* E, const int32_t* mat, void* temp, size_t temp_bytes, uint32_t* keys,
                  uint32_t* keys_out, int32_t* idx, int32_t* idx_sorted, cudaStream_t st);
void mcwk_xs_lookup_staged(const McwLibraryDev& L, int64_t n, const double* E, const int32_t* mat, const int32_t* order,
                           double* out, bool plan, cudaStream_t st);
size_t mcwk_sort_temp_bytes(int64_t n);
void mcwk_compact_status(void* temp, size_t temp_bytes, const int8_t* status, int64_t n, int want, int32_t* queue,
                         unsigned long long* count, cudaStream_t st);
size_t mcwk_compact_temp_bytes(int64_t n);
int mcwk_track_blocks_per_sm(int type, bool fused);
void mcwk_round_reset(McwCounters* c, int consumed, bool fallback_served, cudaStream_t st);
void mcwk_track(const McwParams& P, const int32_t* queue, const unsigned long long* count, int64_t upper, int max_blocks,
                int32_t* q_fly_next, unsigned long long* cnt_next, int32_t* q_collide, int32_t* q_fallback,
                int drain_lanes, bool fused, int collide_min, cudaStream_t st);
void mcwk_collide(const McwParams& P, const int32_t* queue, const unsigned long long* count, int64_t upper,
                  int32_t* q_fly_next, unsigned long long* cnt_next, uint32_t* keys, int32_t* hist, uint32_t* bins,
                  cudaStream_t st);
void mcwk_fallback(const McwParams& P, const int32_t* queue, const unsigned long long* count, int n_blocks,
                   int32_t* q_fly_next, unsigned long long* cnt_next, int32_t* q_collide, cudaStream_t st);
void mcwk_trace_rays(const McwParams& P, int64_t n, const double* start, const double* dirs, const int32_t* start_elem,
                     const int32_t* start_side, const double* max_distance, double* out_distance, int32_t* out_segments,
                     uint64_t* out_hash, int32_t* out_status, cudaStream_t st);
void mcwk_locate_points(const McwParams& P, int64_t n, const double* xyz, const int64_t* ray_id, int32_t* elem,
                        cudaStream_t st);
void mcwk_fill_records(const McwParticlesDev& p, int64_t n, mcw_history_record* rec, cudaStream_t st);
void mcwk_tally_accumulate(double* value, double* sum, double* sum_sq, int64_t n, double norm, cudaStream_t st);
size_t mcwk_scan_temp_bytes(int64_t n);
void mcwk_exclusive_scan(void* temp, size_t temp_bytes, const int64_t* in, int64_t* out, int64_t n, cudaStream_t st);
void mcwk_reduce_sum(void* temp, size_t temp_bytes, const double* in, double* out, int64_t n, cudaStream_t st);
void mcwk_bank_sort(void* temp, size_t temp_bytes, const McwParticlesDev& p, int64_t n_particles, int64_t work_offset,
                    const mcw_source_site* bank, int64_t n_bank, int64_t* scratch_counts, int64_t* scratch_scan,
                    mcw_source_site* sorted, cudaStream_t st);
void mcwk_sample_bank(void* temp, size_t temp_bytes, const mcw_source_site* sorted, int64_t n_bank, int64_t global_start,
                      int64_t total, int64_t n_particles, uint64_t seed0, int64_t* counts, int64_t* offsets,
                      mcw_source_site* out, int64_t cap, cudaStream_t st);
void mcwk_pad_from_bank_end(const mcw_source_site* sorted, int64_t n_bank, mcw_source_site* out, int64_t have,
                            int64_t needed, cudaStream_t st);
uint64_t mcwk_host_init_seed(int64_t id, int offset, int64_t master_seed);
void mcwk_test_division(int64_t n, const double* a, const double* b, double* q_fast, double* q_plain, int32_t* ok, cudaStream_t st);
int mcwk_pump_blocks_per_sm(int type);
void mcwk_pump(const McwParams& P, const McwPump& Q, void* ctl, size_t ctl_bytes, const int32_t* q_fly,
               const unsigned long long* n_fly, const int32_t* q_fallback, const unsigned long long* n_fallback, int blocks,
               int n_track_warps, int patience, cudaStream_t st);
