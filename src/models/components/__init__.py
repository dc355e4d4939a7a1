"""Reference import paths for the hot-path components, backed by the B200 kernels."""
