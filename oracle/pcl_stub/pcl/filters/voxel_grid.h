// stand-in for <pcl/filters/voxel_grid.h> (PCL is not installed here): everything lives in stub_core.h
#pragma once
#include "../stub_core.h"
