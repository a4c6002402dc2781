"""Peer-to-peer plane (CPU side): wire protocol, DHT facade, routing, replication, identity, hardening."""
