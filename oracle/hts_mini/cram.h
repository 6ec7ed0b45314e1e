/* TEST INFRASTRUCTURE ONLY: see sam.h of this directory (everything lives there). */
#include "sam.h"
