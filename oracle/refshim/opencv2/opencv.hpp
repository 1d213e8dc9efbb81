// stand-in: see ../cvstub.h
#include "../cvstub.h"
