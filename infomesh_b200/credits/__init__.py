"""Contribution credits: ledger, tiers, farming detection, verification, cross-node sync (CPU plane)."""
