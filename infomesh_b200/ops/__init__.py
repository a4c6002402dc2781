"""Python entry points of the hand-written sm_100a kernels (one module per kernel family)."""
