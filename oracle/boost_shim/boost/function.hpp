// Shim: the reference includes <boost/function.hpp> but uses nothing from it on the oracle path.
#pragma once
