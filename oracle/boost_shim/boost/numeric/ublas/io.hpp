// Shim: stream output of uBLAS containers is unused on the oracle path.
#pragma once
