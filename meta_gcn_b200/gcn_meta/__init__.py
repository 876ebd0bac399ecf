"""Host-side mirror of the reference package ``gcn_meta`` (src/gcn_meta) for the hot path only."""
