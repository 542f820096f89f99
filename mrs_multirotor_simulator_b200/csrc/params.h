// params.h — host-side parameter derivation (params.cpp)
#pragma once
#include "internal.h"
void mrsb_derive(const mrsb_model_params& mp, const mrsb_controller_params& cp, DevParams* out);
void mrsb_mixer_allocation(const mrsb_model_params& mp, double mix[MRSB_MAX_MOTORS][4]);
