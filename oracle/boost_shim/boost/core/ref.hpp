// Shim: the reference includes <boost/core/ref.hpp> but uses nothing from it on the oracle path.
#pragma once
