#include "mdpp_oracle.h"
