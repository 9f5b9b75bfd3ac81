// stand-in for <pcl/registration/correspondence_rejection_sample_consensus.h> (PCL is not installed here): everything lives in stub_core.h
#pragma once
#include "../stub_core.h"
