#pragma once
#include "open3d/core/Tensor.h"
