#pragma once
#include "scanner/api/kernel.h"
