"""Trust & integrity: Merkle trees, attestations, audits, scoring, reputation, threat detection, DMCA, GDPR."""
