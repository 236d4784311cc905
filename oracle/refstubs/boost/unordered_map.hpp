#pragma once
#include <unordered_map>
namespace boost { using std::unordered_map; }
