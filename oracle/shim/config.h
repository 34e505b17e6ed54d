#pragma once
#define PACKAGE_VERSION "oracle"
