#pragma once
#include "rlog.h"
