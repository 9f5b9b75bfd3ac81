// stand-in for <pcl/features/normal_3d_omp.h> (PCL is not installed here): everything lives in stub_core.h
#pragma once
#include "../stub_core.h"
