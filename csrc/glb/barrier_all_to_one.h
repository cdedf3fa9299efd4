// See barrier_all_to_all.h (both old-style barriers live there).
#pragma once
#include "glb/barrier_all_to_all.h"
