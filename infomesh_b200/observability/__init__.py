"""Metrics, tracing and dashboards."""
