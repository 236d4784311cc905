#pragma once
#define BOOST_VERSION 105500
