// comm.h — thin NCCL wrapper (dlopen at run time; no link-time dependency).
#pragma once
#include "../../include/open3d_b200.h"
