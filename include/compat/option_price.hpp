// compat/option_price.hpp -- the reference's inc/option_price.hpp is empty (:1-6); north_star names
// it as part of the call surface, so it is the home of the new C-ABI declarations.
#pragma once
#include "../mcb200.h"
