// pulls in the reference's own include/preprocess.h with its library includes neutralised (see oracle/pre_stub/pre_stub.h)
#pragma once
#ifndef COMMON_INCLUDE_H
#define COMMON_INCLUDE_H
#endif
#ifndef VELODYNE_CAPTURE
#define VELODYNE_CAPTURE
#endif
#include "pre_stub/pre_stub.h"
#include <algorithm>
#include <preprocess.h>   // -I/root/reference/include
