#pragma once  // stand-in, see nanobind.h
