// stand-in for <pcl/kdtree/kdtree_flann.h> (PCL is not installed here): everything lives in stub_core.h
#pragma once
#include "../stub_core.h"
