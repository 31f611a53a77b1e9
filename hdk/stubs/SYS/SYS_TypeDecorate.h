#pragma once
/* stub */
