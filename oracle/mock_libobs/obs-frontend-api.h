/* mock libobs header (test infrastructure): everything lives in obs-module.h */
#pragma once
#include <obs-module.h>
