#pragma once
#include "../shim_msgs.h"
