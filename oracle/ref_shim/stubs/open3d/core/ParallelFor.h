#pragma once
#include "open3d/core/Indexer.h"
