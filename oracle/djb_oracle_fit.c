/* placeholder, filled in below */
#include "djb_oracle.h"
