// MINIMAL STAND-IN (compile-only check): nothing of this header is used directly by the shim.
