// TEST INFRASTRUCTURE ONLY (oracle build). See format.h.
#pragma once
#include <chrono>
#include "format.h"
