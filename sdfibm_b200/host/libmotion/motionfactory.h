// motionfactory.h — the reference keeps MotionFactory in a header of its own (src/libmotion/motionfactory.h); here it lives in
// imotion.h.  This header exists so that code written against the reference's layout keeps compiling.
#pragma once
#include "imotion.h"
