"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL) rendezvous, symmetric heap, TP layers."""
