// z-score kernels (allele-depth class tally, keep mask, expected/variance moments)
#pragma once
#include "wgs_kernels.cuh"
