// shared_api.hpp — Weed's C shared-library API on the CUDA device.
// Same entry points, argument meaning and error behaviour as the reference's include/shared_api.hpp:34-60 /
// src/shared_api.cpp:79-449 (integer module ids, get_error() codes: 0 ok, 1 module error, 2 invalid argument), so a
// `weed_loader`-style ctypes client written against the reference's libweed_shared binds to libweed_b200.so unchanged.
#pragma once
#include <stddef.h>

typedef unsigned long long uintw;
typedef long long intw;

#ifdef __cplusplus
extern "C" {
#endif
int get_error(const uintw mid);
uintw load_module(const char *f);
void save_module(uintw mid, const char *f);
void free_module(uintw mid);
void forward(uintw mid, uintw dtype, uintw n, uintw *shape, double *d);
void forward_int(uintw mid, uintw dtype, uintw n, uintw *shape, intw *d);
uintw get_result_index_count(uintw mid);
void get_result_dims(uintw mid, uintw *shape, uintw *stride);
uintw get_result_size(uintw mid);
uintw get_result_offset(uintw mid);
uintw get_result_type(uintw mid);
void get_result(uintw mid, double *d);
void train_step(uintw mid, uintw n, uintw *shape, intw *input_ids, uintw n_target, intw *target_ids, double learning_rate);
void reset_kv_cache(uintw mid);
void set_max_kv_seq_len(uintw mid, uintw m);
#ifdef __cplusplus
}
#endif
