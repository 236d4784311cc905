#pragma once
#include <unordered_set>
namespace boost { using std::unordered_set; }
