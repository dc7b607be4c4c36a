// nmf_step_all.cuh — the step kernel source in both precisions (see nmf_step_common.cuh).
#pragma once
#include "nmf_step_common.cuh"

#define real float
#define NMF_NS f32
#include "nmf_step.cuh"
#undef real
#undef NMF_NS

#define real double
#define NMF_NS f64
#include "nmf_step.cuh"
#undef real
#undef NMF_NS
