"""Local HTTP admin API and API extensions (OpenAPI spec, rate limiting, API keys, shell completion)."""
