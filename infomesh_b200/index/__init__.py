"""Index layer: SQLite/FTS5 document store, GPU shards, ranking, link graph, snapshots, importers."""
