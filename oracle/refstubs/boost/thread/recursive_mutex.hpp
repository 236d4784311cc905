#pragma once
#include <boost/thread/thread.hpp>
