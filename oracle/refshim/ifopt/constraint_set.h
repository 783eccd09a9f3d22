// stand-in: see composite.h
#include <ifopt/composite.h>
