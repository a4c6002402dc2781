"""Dashboard tab panes."""
