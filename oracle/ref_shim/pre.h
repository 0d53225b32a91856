#pragma once
#ifdef __CUDACC__
#include <thrust/sort.h>
#include <thrust/execution_policy.h>
#endif
