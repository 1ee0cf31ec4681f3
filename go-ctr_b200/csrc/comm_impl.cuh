// comm_impl.cuh — included by engine.cu after ctr_handle and step_core are defined.
#pragma once

static int comm_allreduce_grads(ctr_handle* h) { return set_err(h, CTR_ESTATE, "multi-GPU path not initialised"); }
static int comm_train_step(ctr_handle* h, const int32_t*, const int32_t*, const int32_t*, const float*, int32_t) { return set_err(h, CTR_ESTATE, "multi-GPU path not initialised"); }
static int comm_predict(ctr_handle* h, const int32_t*, const int32_t*, const int32_t*, int32_t, float*) { return set_err(h, CTR_ESTATE, "multi-GPU path not initialised"); }
static int comm_unique_id(void*, int32_t*) { return CTR_ESTATE; }
static int comm_init(ctr_handle* h, const void*, int32_t) { return set_err(h, CTR_ESTATE, "multi-GPU path not built yet"); }
static void comm_destroy(ctr_handle*) {}
