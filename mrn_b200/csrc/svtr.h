#pragma once
#include "../../include/mrn_b200.h"
