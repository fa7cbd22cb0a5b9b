// Stub for RVI/factor/marginalization_factor.h: common_function.h only needs the type name
// (mea_t holds a MarginalizationInfo*), plus the <cstdint>/<cstring> the real header drags in.
#pragma once
#include <cstdint>
#include <cstring>
class MarginalizationInfo;
