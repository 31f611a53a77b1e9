#pragma once
#include <UT/UT_VectorTypes.h>
