"""Persistent server-side state (analytics, webhooks, sessions, history, presets)."""
