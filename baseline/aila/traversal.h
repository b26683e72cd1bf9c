/* Stand-in for the generated tools/common/traversal.h the comparator includes: the PODs it uses are
 * declared in this repository's C ABI header with the reference's layouts. */
#pragma once
#include "rodent_b200.h"
