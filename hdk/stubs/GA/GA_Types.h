#pragma once
#include <SYS/SYS_Types.h>
typedef exint GA_Size;
typedef exint GA_Offset;
