#include "../mock.h"
