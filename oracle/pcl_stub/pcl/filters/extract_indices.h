// stand-in for <pcl/filters/extract_indices.h> (PCL is not installed here): everything lives in stub_core.h
#pragma once
#include "../stub_core.h"
