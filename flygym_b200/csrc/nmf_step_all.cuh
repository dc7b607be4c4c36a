// nmf_step_all.cuh — the step kernel source in both precisions (see nmf_step_common.cuh).
#pragma once
#include "nmf_step_common.cuh"
#include "nmf_tree_layout.h"

#define real float
#define NMF_NS f32
#include "nmf_step.cuh"
#include "nmf_tree.cuh"
#undef real
#undef NMF_NS

#define real double
#define NMF_NS f64
#include "nmf_step.cuh"
#include "nmf_tree.cuh"
#undef real
#undef NMF_NS
