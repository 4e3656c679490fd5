"""Differentiable collectives.

The adjoint table is the reference's (``epl/communicators/nccl_ops.py:37-124``):
all-reduce <-> all-reduce, all-gather <-> reduce-scatter, reduce-scatter <->
all-gather, reduce <-> broadcast, all-to-all <-> all-to-all (v: with swapped
counts); broadcast of parameters and gatherv are not differentiated.
"""
from __future__ import annotations

import torch


class _AllReduce(torch.autograd.Function):
  @staticmethod
  def forward(ctx, t, comm, op):
    ctx.comm, ctx.op = comm, op
    return comm.primary.all_reduce(t.contiguous().clone(), op)

  @staticmethod
  def backward(ctx, g):
    if ctx.op != "sum":
      raise RuntimeError("only sum all-reduce is differentiable")
    return ctx.comm.primary.all_reduce(g.contiguous().clone(), "sum"), None, None


class _AllGather(torch.autograd.Function):
  @staticmethod
  def forward(ctx, t, comm):
    ctx.comm = comm
    return comm.primary.all_gather(t)

  @staticmethod
  def backward(ctx, g):
    return ctx.comm.primary.reduce_scatter(g.contiguous(), "sum"), None


class _ReduceScatter(torch.autograd.Function):
  @staticmethod
  def forward(ctx, t, comm):
    ctx.comm = comm
    return comm.primary.reduce_scatter(t, "sum")

  @staticmethod
  def backward(ctx, g):
    return ctx.comm.primary.all_gather(g.contiguous()), None


class _AllToAll(torch.autograd.Function):
  @staticmethod
  def forward(ctx, t, comm):
    ctx.comm = comm
    return comm.alltoall(t)

  @staticmethod
  def backward(ctx, g):
    return ctx.comm.alltoall(g.contiguous()), None


class _AllToAllV(torch.autograd.Function):
  @staticmethod
  def forward(ctx, t, send_counts, comm):
    out, recv_counts = comm.alltoallv(t, send_counts)
    ctx.comm, ctx.recv_counts = comm, recv_counts
    ctx.mark_non_differentiable(recv_counts)
    return out, recv_counts

  @staticmethod
  def backward(ctx, g, _):
    back, _ = ctx.comm.alltoallv(g.contiguous(), ctx.recv_counts)
    return back, None, None


class _Reduce(torch.autograd.Function):
  @staticmethod
  def forward(ctx, t, comm, root):
    ctx.comm, ctx.root = comm, root
    return comm.primary.reduce(t.contiguous().clone(), root, "sum")

  @staticmethod
  def backward(ctx, g):
    return ctx.comm.primary.broadcast(g.contiguous().clone(), ctx.root), None, None


class _CopyToGroup(torch.autograd.Function):
  """Identity forward, all-reduce backward (input of a column-parallel layer)."""
  @staticmethod
  def forward(ctx, t, comm):
    ctx.comm = comm
    return t

  @staticmethod
  def backward(ctx, g):
    return ctx.comm.primary.all_reduce(g.contiguous().clone(), "sum"), None


def all_reduce(t, comm, op="sum"): return _AllReduce.apply(t, comm, op) if comm.size > 1 else t
def all_gather(t, comm): return _AllGather.apply(t, comm) if comm.size > 1 else t
def reduce_scatter(t, comm): return _ReduceScatter.apply(t, comm) if comm.size > 1 else t
def all_to_all(t, comm): return _AllToAll.apply(t, comm) if comm.size > 1 else t
def all_to_allv(t, send_counts, comm): return _AllToAllV.apply(t, send_counts, comm)
def reduce(t, comm, root=0): return _Reduce.apply(t, comm, root) if comm.size > 1 else t
def copy_to_group(t, comm): return _CopyToGroup.apply(t, comm) if comm.size > 1 else t
