#pragma once
/* stub */
class GU_Detail;
