"""Small shared utilities (tokenisers, timers, logging shim)."""
