/* placeholder translation unit: compute_quotient_polys restatement lands here (see p2oracle.h). */
#include "p2oracle.h"
