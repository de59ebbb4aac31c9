// oracle stub, see ImfRgbaFile.h
#include "ImfRgbaFile.h"
