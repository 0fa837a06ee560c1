// TEST INFRASTRUCTURE ONLY (oracle build). See ../spdlog.h.
#pragma once
#include "../spdlog.h"
