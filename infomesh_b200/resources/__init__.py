"""Resource governance: profiles, the load governor, preflight checks, port reachability."""
