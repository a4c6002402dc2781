"""Terminal dashboard: Textual TUI (six tabs) and a Rich one-shot text report."""
