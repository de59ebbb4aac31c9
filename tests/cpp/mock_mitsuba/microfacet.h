#include "mitsuba/mock.h"
