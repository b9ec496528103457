// Backend.Cuda.fs — managed shim: Backend<'T>, ISigmaDiffDataBuffer<'T> and IBackendProvider over
// libdiffsharp_cuda.so (P/Invoke), to be compiled into the CONSUMER of Sigma.DiffSharp next to (or
// instead of) its OpenBLAS backend.  Covers 'T = float32 AND 'T = float (Config.fs:43-54 needs both).
//
// STATUS: NEVER COMPILED OR RUN.  This image has no .NET / Mono / F# toolchain.  The file documents the
// binding a maintainer would add; its twin that IS executed by the test-suite is the Python ctypes
// layer (sigma/diffsharp_b200/_lib.py, device.py, backend.py), which makes the same calls in the same
// order: ONE native call per Backend<'T> member.  Nothing is decided on this side — the deferred
// evaluation that turns the AD layer's call chains into single kernels (lazy transposes, bias broadcast
// + activation, the tanh / sigmoid / ReLU adjoint chains, zero adjoints) lives inside the native
// library (csrc/dsc_lazy.cu), so this shim inherits it.
namespace DiffSharp.Backend.Cuda

open System
open System.Runtime.InteropServices
open System.Runtime.CompilerServices
open DiffSharp.Util
open DiffSharp.Backend
open DiffSharp.Config

/// MapOp is a plain union (Backend.fs:42-70), not an enum: the op-code that crosses the ABI is its
/// position in declaration order, spelled out here (include/diffsharp_cuda.h `enum dsc_mapop`).
module OpCode =
    let ofMapOp (op : MapOp) : int =
        match op with
        | MapOp.Mul -> 0 | MapOp.Div -> 1 | MapOp.Add -> 2 | MapOp.Sub -> 3
        | MapOp.Pow_Ab -> 4 | MapOp.Pow_Ba -> 5 | MapOp.Atan2_Ab -> 6 | MapOp.Atan2_Ba -> 7
        | MapOp.Log -> 8 | MapOp.Log10 -> 9 | MapOp.Exp -> 10 | MapOp.Sin -> 11 | MapOp.Cos -> 12
        | MapOp.Tan -> 13 | MapOp.Sqrt -> 14 | MapOp.Sinh -> 15 | MapOp.Cosh -> 16 | MapOp.Tanh -> 17
        | MapOp.Asin -> 18 | MapOp.Acos -> 19 | MapOp.Atan -> 20 | MapOp.Abs -> 21 | MapOp.Sign -> 22
        | MapOp.Floor -> 23 | MapOp.Ceiling -> 24 | MapOp.Round -> 25 | MapOp.ReL -> 26 | MapOp.Sigmoid -> 27

module Native =
    [<Literal>]
    let Lib = "diffsharp_cuda" // libdiffsharp_cuda.so on the library path

    // ---- untyped surface -------------------------------------------------------------------------
    [<DllImport(Lib)>] extern int dsc_create(int device, uint32 flags, nativeint& ctx)
    [<DllImport(Lib)>] extern int dsc_destroy(nativeint ctx)
    [<DllImport(Lib)>] extern int dsc_sync(nativeint ctx)
    [<DllImport(Lib)>] extern nativeint dsc_last_error()
    [<DllImport(Lib)>] extern int dsc_buf_alloc(nativeint ctx, int dtype, int n, uint64& h)
    [<DllImport(Lib)>] extern int dsc_buf_alloc_fill(nativeint ctx, int dtype, int n, double v, uint64& h)
    [<DllImport(Lib)>] extern int dsc_buf_free(uint64 h)
    [<DllImport(Lib)>] extern int dsc_buf_upload(uint64 h, int off, nativeint host, int n)     // pinned GCHandle address of a 'T[]
    [<DllImport(Lib)>] extern int dsc_buf_download(uint64 h, int off, nativeint host, int n)
    [<DllImport(Lib)>] extern int dsc_buf_view(uint64 h, int off, int n, uint64& out)
    [<DllImport(Lib)>] extern int dsc_buf_copy(uint64 h, uint64& out)
    [<DllImport(Lib)>] extern int dsc_buf_gather2d(uint64 h, int rows, int cols, int r0, int r1, int c0, int c1, uint64& out)

    // ---- typed surface: dsc_s* (float32) and dsc_d* (float), same shapes ---------------------------
    [<DllImport(Lib)>] extern int dsc_sgemm(nativeint ctx, int ta, int tb, int m, int n, int k, uint64 A, uint64 B, uint64 bias, uint64& C)
    [<DllImport(Lib)>] extern int dsc_dgemm(nativeint ctx, int ta, int tb, int m, int n, int k, uint64 A, uint64 B, uint64 bias, uint64& C)
    [<DllImport(Lib)>] extern int dsc_sgemv(nativeint ctx, int trans, int m, int n, uint64 A, uint64 x, uint64 y, uint64& out)
    [<DllImport(Lib)>] extern int dsc_dgemv(nativeint ctx, int trans, int m, int n, uint64 A, uint64 x, uint64 y, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sger(nativeint ctx, uint64 x, uint64 y, uint64& out)
    [<DllImport(Lib)>] extern int dsc_dger(nativeint ctx, uint64 x, uint64 y, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sdot(nativeint ctx, uint64 x, uint64 y, float32& out)
    [<DllImport(Lib)>] extern int dsc_ddot(nativeint ctx, uint64 x, uint64 y, float& out)
    [<DllImport(Lib)>] extern int dsc_sasum(nativeint ctx, uint64 x, float32& out)
    [<DllImport(Lib)>] extern int dsc_dasum(nativeint ctx, uint64 x, float& out)
    [<DllImport(Lib)>] extern int dsc_snrm2(nativeint ctx, uint64 x, float32& out)
    [<DllImport(Lib)>] extern int dsc_dnrm2(nativeint ctx, uint64 x, float& out)
    [<DllImport(Lib)>] extern int dsc_samax(nativeint ctx, uint64 x, float32& out)
    [<DllImport(Lib)>] extern int dsc_damax(nativeint ctx, uint64 x, float& out)
    [<DllImport(Lib)>] extern int dsc_ssum(nativeint ctx, uint64 x, float32& out)
    [<DllImport(Lib)>] extern int dsc_dsum(nativeint ctx, uint64 x, float& out)
    [<DllImport(Lib)>] extern int dsc_sargmax(nativeint ctx, uint64 x, int& out)
    [<DllImport(Lib)>] extern int dsc_dargmax(nativeint ctx, uint64 x, int& out)
    [<DllImport(Lib)>] extern int dsc_sargmin(nativeint ctx, uint64 x, int& out)
    [<DllImport(Lib)>] extern int dsc_dargmin(nativeint ctx, uint64 x, int& out)
    [<DllImport(Lib)>] extern int dsc_smap(nativeint ctx, int op, uint64 x, uint64& y)
    [<DllImport(Lib)>] extern int dsc_dmap(nativeint ctx, int op, uint64 x, uint64& y)
    [<DllImport(Lib)>] extern int dsc_smap_s(nativeint ctx, int op, float32 s, uint64 x, uint64& y)
    [<DllImport(Lib)>] extern int dsc_dmap_s(nativeint ctx, int op, float s, uint64 x, uint64& y)
    [<DllImport(Lib)>] extern int dsc_smap2(nativeint ctx, int op, uint64 a, uint64 b, uint64& y)
    [<DllImport(Lib)>] extern int dsc_dmap2(nativeint ctx, int op, uint64 a, uint64 b, uint64& y)
    [<DllImport(Lib)>] extern int dsc_sscalar_op(nativeint ctx, int kind, float32 s, uint64 x, uint64& y)
    [<DllImport(Lib)>] extern int dsc_dscalar_op(nativeint ctx, int kind, float s, uint64 x, uint64& y)
    [<DllImport(Lib)>] extern int dsc_sadd_inplace(nativeint ctx, uint64 dst, uint64 src)
    [<DllImport(Lib)>] extern int dsc_dadd_inplace(nativeint ctx, uint64 dst, uint64 src)
    [<DllImport(Lib)>] extern int dsc_sadd_inplace_range(nativeint ctx, uint64 src, int soff, uint64 dst, int doff, int n)
    [<DllImport(Lib)>] extern int dsc_dadd_inplace_range(nativeint ctx, uint64 src, int soff, uint64 dst, int doff, int n)
    [<DllImport(Lib)>] extern int dsc_stranspose(nativeint ctx, int m, int n, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_dtranspose(nativeint ctx, int m, int n, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_spermute(nativeint ctx, int rank, int64[] shape, int[] perm, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_dpermute(nativeint ctx, int rank, int64[] shape, int[] perm, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_srepeat_rows(nativeint ctx, int m, uint64 v, uint64& out)
    [<DllImport(Lib)>] extern int dsc_drepeat_rows(nativeint ctx, int m, uint64 v, uint64& out)
    [<DllImport(Lib)>] extern int dsc_srepeat_cols(nativeint ctx, int n, uint64 v, uint64& out)
    [<DllImport(Lib)>] extern int dsc_drepeat_cols(nativeint ctx, int n, uint64 v, uint64& out)
    [<DllImport(Lib)>] extern int dsc_scolsum_inplace(nativeint ctx, int m, int n, uint64 M, uint64 v)
    [<DllImport(Lib)>] extern int dsc_dcolsum_inplace(nativeint ctx, int m, int n, uint64 M, uint64 v)
    [<DllImport(Lib)>] extern int dsc_sadd_v_mcols(nativeint ctx, int m, int n, uint64 v, uint64 M, uint64& out)
    [<DllImport(Lib)>] extern int dsc_dadd_v_mcols(nativeint ctx, int m, int n, uint64 v, uint64 M, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sdiag(nativeint ctx, int m, int n, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_ddiag(nativeint ctx, int m, int n, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sgesv(nativeint ctx, int n, uint64 A, uint64 b, uint64& x)
    [<DllImport(Lib)>] extern int dsc_dgesv(nativeint ctx, int n, uint64 A, uint64 b, uint64& x)
    [<DllImport(Lib)>] extern int dsc_ssysv(nativeint ctx, int n, uint64 A, uint64 b, uint64& x)
    [<DllImport(Lib)>] extern int dsc_dsysv(nativeint ctx, int n, uint64 A, uint64 b, uint64& x)
    [<DllImport(Lib)>] extern int dsc_sgetri(nativeint ctx, int n, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_dgetri(nativeint ctx, int n, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sdet(nativeint ctx, int n, uint64 A, float32& out)
    [<DllImport(Lib)>] extern int dsc_ddet(nativeint ctx, int n, uint64 A, float& out)

    let DSC_E_SHAPE, DSC_E_UNSUPPORTED_OP, DSC_E_SINGULAR = -2, -6, -7
    /// status -> .NET exception, mirroring the reference's own error style (Util.fs:117-131)
    let check (status : int) =
        if status <> 0 then
            let msg = Marshal.PtrToStringAnsi(dsc_last_error())
            match status with
            | s when s = DSC_E_SHAPE -> invalidArg "" msg
            | s when s = DSC_E_UNSUPPORTED_OP -> raise (NotSupportedException msg) // closures are never executed
            | _ -> failwithf "libdiffsharp_cuda: %s" msg
    /// DSC_E_SINGULAR -> None (Backend.fs:102-103,125-126: the caller raises invalidArg, AD.Float32.fs:2816-2818)
    let optional (status : int) (value : unit -> 'a) : 'a option =
        if status = DSC_E_SINGULAR then None
        else check status; Some(value ())

/// The typed entry points of one precision, so that ONE generic Backend<'T> serves float32 and float.
type NativeOps<'T> =
    { dtype : int
      gemm : nativeint * int * int * int * int * int * uint64 * uint64 * uint64 -> uint64
      gemv : nativeint * int * int * int * uint64 * uint64 * uint64 -> uint64
      ger : nativeint * uint64 * uint64 -> uint64
      dot : nativeint * uint64 * uint64 -> 'T
      asum : nativeint * uint64 -> 'T
      nrm2 : nativeint * uint64 -> 'T
      amax : nativeint * uint64 -> 'T
      sum : nativeint * uint64 -> 'T
      argmax : nativeint * uint64 -> int
      argmin : nativeint * uint64 -> int
      map : nativeint * int * uint64 -> uint64
      map_s : nativeint * int * 'T * uint64 -> uint64
      map2 : nativeint * int * uint64 * uint64 -> uint64
      scalar_op : nativeint * int * 'T * uint64 -> uint64
      add_inplace : nativeint * uint64 * uint64 -> unit
      add_inplace_range : nativeint * uint64 * int * uint64 * int * int -> unit
      transpose : nativeint * int * int * uint64 -> uint64
      permute : nativeint * int64[] * int[] * uint64 -> uint64
      repeat_rows : nativeint * int * uint64 -> uint64
      repeat_cols : nativeint * int * uint64 -> uint64
      colsum_inplace : nativeint * int * int * uint64 * uint64 -> unit
      add_v_mcols : nativeint * int * int * uint64 * uint64 -> uint64
      diag : nativeint * int * int * uint64 -> uint64
      gesv : nativeint * int * uint64 * uint64 -> uint64 option
      sysv : nativeint * int * uint64 * uint64 -> uint64 option
      getri : nativeint * int * uint64 -> uint64 option
      det : nativeint * int * uint64 -> 'T option
      toDouble : 'T -> double }

module Ops =
    open Native
    let inline private out (f : byref<uint64> -> int) = let mutable h = 0UL in check (f &h); h
    let float32Ops : NativeOps<float32> =
        { dtype = 0
          gemm = fun (c, ta, tb, m, n, k, a, b, bias) -> let mutable h = 0UL in check (dsc_sgemm(c, ta, tb, m, n, k, a, b, bias, &h)); h
          gemv = fun (c, t, m, n, a, x, y) -> let mutable h = 0UL in check (dsc_sgemv(c, t, m, n, a, x, y, &h)); h
          ger = fun (c, x, y) -> let mutable h = 0UL in check (dsc_sger(c, x, y, &h)); h
          dot = fun (c, x, y) -> let mutable r = 0.f in check (dsc_sdot(c, x, y, &r)); r
          asum = fun (c, x) -> let mutable r = 0.f in check (dsc_sasum(c, x, &r)); r
          nrm2 = fun (c, x) -> let mutable r = 0.f in check (dsc_snrm2(c, x, &r)); r
          amax = fun (c, x) -> let mutable r = 0.f in check (dsc_samax(c, x, &r)); r
          sum = fun (c, x) -> let mutable r = 0.f in check (dsc_ssum(c, x, &r)); r
          argmax = fun (c, x) -> let mutable r = 0 in check (dsc_sargmax(c, x, &r)); r
          argmin = fun (c, x) -> let mutable r = 0 in check (dsc_sargmin(c, x, &r)); r
          map = fun (c, op, x) -> let mutable h = 0UL in check (dsc_smap(c, op, x, &h)); h
          map_s = fun (c, op, s, x) -> let mutable h = 0UL in check (dsc_smap_s(c, op, s, x, &h)); h
          map2 = fun (c, op, a, b) -> let mutable h = 0UL in check (dsc_smap2(c, op, a, b, &h)); h
          scalar_op = fun (c, kind, s, x) -> let mutable h = 0UL in check (dsc_sscalar_op(c, kind, s, x, &h)); h
          add_inplace = fun (c, d, s) -> check (dsc_sadd_inplace(c, d, s))
          add_inplace_range = fun (c, s, so, d, dof, n) -> check (dsc_sadd_inplace_range(c, s, so, d, dof, n))
          transpose = fun (c, m, n, a) -> let mutable h = 0UL in check (dsc_stranspose(c, m, n, a, &h)); h
          permute = fun (c, shape, perm, a) -> let mutable h = 0UL in check (dsc_spermute(c, shape.Length, shape, perm, a, &h)); h
          repeat_rows = fun (c, m, v) -> let mutable h = 0UL in check (dsc_srepeat_rows(c, m, v, &h)); h
          repeat_cols = fun (c, n, v) -> let mutable h = 0UL in check (dsc_srepeat_cols(c, n, v, &h)); h
          colsum_inplace = fun (c, m, n, mh, v) -> check (dsc_scolsum_inplace(c, m, n, mh, v))
          add_v_mcols = fun (c, m, n, v, mh) -> let mutable h = 0UL in check (dsc_sadd_v_mcols(c, m, n, v, mh, &h)); h
          diag = fun (c, m, n, a) -> let mutable h = 0UL in check (dsc_sdiag(c, m, n, a, &h)); h
          gesv = fun (c, n, a, b) -> let mutable h = 0UL in optional (dsc_sgesv(c, n, a, b, &h)) (fun () -> h)
          sysv = fun (c, n, a, b) -> let mutable h = 0UL in optional (dsc_ssysv(c, n, a, b, &h)) (fun () -> h)
          getri = fun (c, n, a) -> let mutable h = 0UL in optional (dsc_sgetri(c, n, a, &h)) (fun () -> h)
          det = fun (c, n, a) -> let mutable r = 0.f in optional (dsc_sdet(c, n, a, &r)) (fun () -> r)
          toDouble = double }
    /// the float64 twin: the same record over the dsc_d* entry points
    let float64Ops : NativeOps<float> =
        { dtype = 1
          gemm = fun (c, ta, tb, m, n, k, a, b, bias) -> let mutable h = 0UL in check (dsc_dgemm(c, ta, tb, m, n, k, a, b, bias, &h)); h
          gemv = fun (c, t, m, n, a, x, y) -> let mutable h = 0UL in check (dsc_dgemv(c, t, m, n, a, x, y, &h)); h
          ger = fun (c, x, y) -> let mutable h = 0UL in check (dsc_dger(c, x, y, &h)); h
          dot = fun (c, x, y) -> let mutable r = 0. in check (dsc_ddot(c, x, y, &r)); r
          asum = fun (c, x) -> let mutable r = 0. in check (dsc_dasum(c, x, &r)); r
          nrm2 = fun (c, x) -> let mutable r = 0. in check (dsc_dnrm2(c, x, &r)); r
          amax = fun (c, x) -> let mutable r = 0. in check (dsc_damax(c, x, &r)); r
          sum = fun (c, x) -> let mutable r = 0. in check (dsc_dsum(c, x, &r)); r
          argmax = fun (c, x) -> let mutable r = 0 in check (dsc_dargmax(c, x, &r)); r
          argmin = fun (c, x) -> let mutable r = 0 in check (dsc_dargmin(c, x, &r)); r
          map = fun (c, op, x) -> let mutable h = 0UL in check (dsc_dmap(c, op, x, &h)); h
          map_s = fun (c, op, s, x) -> let mutable h = 0UL in check (dsc_dmap_s(c, op, s, x, &h)); h
          map2 = fun (c, op, a, b) -> let mutable h = 0UL in check (dsc_dmap2(c, op, a, b, &h)); h
          scalar_op = fun (c, kind, s, x) -> let mutable h = 0UL in check (dsc_dscalar_op(c, kind, s, x, &h)); h
          add_inplace = fun (c, d, s) -> check (dsc_dadd_inplace(c, d, s))
          add_inplace_range = fun (c, s, so, d, dof, n) -> check (dsc_dadd_inplace_range(c, s, so, d, dof, n))
          transpose = fun (c, m, n, a) -> let mutable h = 0UL in check (dsc_dtranspose(c, m, n, a, &h)); h
          permute = fun (c, shape, perm, a) -> let mutable h = 0UL in check (dsc_dpermute(c, shape.Length, shape, perm, a, &h)); h
          repeat_rows = fun (c, m, v) -> let mutable h = 0UL in check (dsc_drepeat_rows(c, m, v, &h)); h
          repeat_cols = fun (c, n, v) -> let mutable h = 0UL in check (dsc_drepeat_cols(c, n, v, &h)); h
          colsum_inplace = fun (c, m, n, mh, v) -> check (dsc_dcolsum_inplace(c, m, n, mh, v))
          add_v_mcols = fun (c, m, n, v, mh) -> let mutable h = 0UL in check (dsc_dadd_v_mcols(c, m, n, v, mh, &h)); h
          diag = fun (c, m, n, a) -> let mutable h = 0UL in check (dsc_ddiag(c, m, n, a, &h)); h
          gesv = fun (c, n, a, b) -> let mutable h = 0UL in optional (dsc_dgesv(c, n, a, b, &h)) (fun () -> h)
          sysv = fun (c, n, a, b) -> let mutable h = 0UL in optional (dsc_dsysv(c, n, a, b, &h)) (fun () -> h)
          getri = fun (c, n, a) -> let mutable h = 0UL in optional (dsc_dgetri(c, n, a, &h)) (fun () -> h)
          det = fun (c, n, a) -> let mutable r = 0. in optional (dsc_ddet(c, n, a, &r)) (fun () -> r)
          toDouble = id }

/// Device-resident ISigmaDiffDataBuffer<'T> (Util.fs:133-141).  Owns one dsc_buf_t.
/// `Data` is a cached host mirror of the whole allocation, shared by the views of it; host-side writes into it
/// (`aa.Data.[i] <- ...`) are detected and uploaded by `Sync()` before the next device use; device-side writes call
/// `Invalidate()` (the Python twin: device.py `_flush_host_writes` / `_invalidate_host`).
[<AllowNullLiteral>]
type CudaDataBuffer<'T when 'T : unmanaged and 'T : equality>(ctx : nativeint, handle : uint64, length : int, offset : int, root : CudaDataBuffer<'T>) =
    let mutable mirror : 'T[] = null
    let mutable clean : 'T[] = null
    member _.Handle = handle
    member _.Ctx = ctx
    member d.Owner : CudaDataBuffer<'T> = if isNull root then d else root
    member private d.Pinned (a : 'T[]) (f : nativeint -> int) =
        let g = GCHandle.Alloc(a, GCHandleType.Pinned)
        try Native.check (f (g.AddrOfPinnedObject())) finally g.Free()
    member d.Download() : 'T[] =
        let a : 'T[] = Array.zeroCreate length
        if length > 0 then d.Pinned a (fun p -> Native.dsc_buf_download(handle, 0, p, length))
        a
    member d.Mirror with get () = mirror and set v = mirror <- v
    member d.Clean with get () = clean and set v = clean <- v
    /// upload host-side writes made through `.Data` since the mirror was taken
    member d.Sync() =
        let o = d.Owner
        if not (isNull o.Mirror) && o.Mirror <> o.Clean then
            o.Pinned o.Mirror (fun p -> Native.dsc_buf_upload(o.Handle, 0, p, o.Mirror.Length))
            o.Clean <- Array.copy o.Mirror
    member d.Invalidate() = let o = d.Owner in o.Mirror <- null; o.Clean <- null
    override _.Finalize() = Native.dsc_buf_free handle |> ignore
    interface ISigmaDiffDataBuffer<'T> with
        member _.Length = length
        member _.Offset = offset
        member d.SubData = d.Sync(); d.Download()
        member d.Data =
            let o = d.Owner
            if isNull o.Mirror then
                o.Mirror <- o.Download()
                o.Clean <- Array.copy o.Mirror
            o.Mirror
        member d.GetValues startIndex len =
            d.Sync()
            let mutable h = 0UL
            Native.check (Native.dsc_buf_view(handle, startIndex, len, &h))
            CudaDataBuffer<'T>(ctx, h, len, offset + startIndex, d.Owner) :> ISigmaDiffDataBuffer<'T>
        member d.GetStackedValues rows cols r0 r1 c0 c1 =
            d.Sync()
            let mutable h = 0UL
            Native.check (Native.dsc_buf_gather2d(handle, rows, cols, r0, r1, c0, c1, &h))
            CudaDataBuffer<'T>(ctx, h, (r1 - r0 + 1) * (c1 - c0 + 1), 0, null) :> ISigmaDiffDataBuffer<'T>
        member d.DeepCopy() =
            d.Sync()
            let mutable h = 0UL
            Native.check (Native.dsc_buf_copy(handle, &h))
            CudaDataBuffer<'T>(ctx, h, length, 0, null) :> ISigmaDiffDataBuffer<'T>
        member d.ShallowCopy() = (d :> ISigmaDiffDataBuffer<'T>).GetValues 0 length

/// What a consumer passes as `customInfo : obj` to DNDArray.CustomOp (AD.Float32.fs:1758-1763, 4282): its own forward and
/// adjoint, written against the backend it is handed.
type ICudaCustomOp<'T> =
    abstract Forward : Backend<'T> * ShapedDataBufferView<'T> -> ShapedDataBufferView<'T>
    abstract Backward : Backend<'T> * ShapedDataBufferView<'T> * ShapedDataBufferView<'T> * ShapedDataBufferView<'T> -> ShapedDataBufferView<'T>

/// Backend<'T> over the C ABI.  One P/Invoke per member.
type CudaBackend<'T when 'T : unmanaged and 'T : equality>(ctx : nativeint, ops : NativeOps<'T>, zero : 'T) =
    // CreateZeroArray / CreateValueArray return managed arrays registered here as "n copies of v" so that
    // CreateDataBuffer can fill on the device instead of touching the host pages (SURVEY §7.3).
    let constants = ConditionalWeakTable<'T[], obj>()
    let wrap n (h : uint64) = CudaDataBuffer<'T>(ctx, h, n, 0, null) :> ISigmaDiffDataBuffer<'T>
    let upload (a : 'T[]) =
        let mutable h = 0UL
        Native.check (Native.dsc_buf_alloc(ctx, ops.dtype, a.Length, &h))
        if a.Length > 0 then
            let g = GCHandle.Alloc(a, GCHandleType.Pinned)
            try Native.check (Native.dsc_buf_upload(h, 0, g.AddrOfPinnedObject(), a.Length)) finally g.Free()
        h
    let dev (b : ISigmaDiffDataBuffer<'T>) : CudaDataBuffer<'T> =
        match b with
        | :? CudaDataBuffer<'T> as c -> c.Sync(); c
        | _ -> // foreign buffer (DVector.ZeroN / DNDArray.Zero, AD.Float32.fs:706-707,1833): upload on first touch
            CudaDataBuffer<'T>(ctx, upload b.SubData, b.Length, 0, null)
    let h (b : ISigmaDiffDataBuffer<'T>) = (dev b).Handle
    let view (b : ISigmaDiffDataBuffer<'T>) ([<ParamArray>] shape : int64[]) = ShapedDataBufferView<'T>(b, shape)
    let map2 op (a : ISigmaDiffDataBuffer<'T>) (b : ISigmaDiffDataBuffer<'T>) = wrap (max a.Length b.Length) (ops.map2 (ctx, op, h a, h b))
    let scalarOp kind s (v : ISigmaDiffDataBuffer<'T>) = wrap v.Length (ops.scalar_op (ctx, kind, s, h v))
    interface Backend<'T> with
        member _.CreateDataBuffer a =
            match constants.TryGetValue a with
            | true, (:? 'T as v) ->
                let mutable hh = 0UL
                Native.check (Native.dsc_buf_alloc_fill(ctx, ops.dtype, a.Length, ops.toDouble v, &hh))
                wrap a.Length hh
            | _ -> wrap a.Length (upload a)
        member _.CreateUninitialisedArray n = Array.zeroCreate n
        member _.CreateZeroArray n = let a = Array.zeroCreate n in constants.Add(a, box zero); a
        member _.CreateValueArray(n, v) = let a = Array.create n v in constants.Add(a, box v); a
        member _.Mul_Dot_V_V(a, b) = ops.dot (ctx, h a, h b)
        member _.L1Norm_V a = ops.asum (ctx, h a)
        member _.L2Norm_V a = ops.nrm2 (ctx, h a)
        member _.SupNorm_V a = ops.amax (ctx, h a)
        member _.Sum_V a = ops.sum (ctx, h a)
        member _.Sum_M a = ops.sum (ctx, h a)
        member _.MaxIndex_V a = ops.argmax (ctx, h a)
        member _.MinIndex_V a = ops.argmin (ctx, h a)
        member _.Add_V_V(a, b) = map2 2 a b
        member _.Sub_V_V(a, b) = map2 3 a b
        member _.Add_V_V_InPlace(src, so, dst, dof, n) =
            let d = dev dst
            ops.add_inplace_range (ctx, h src, so, d.Handle, dof, n); d.Invalidate(); dst
        member _.Add_S_V(s, v) = scalarOp 0 s v
        member _.Sub_S_V(s, v) = scalarOp 1 s v
        member _.Sub_V_S(v, s) = scalarOp 2 s v
        member _.Mul_S_V(s, v) = scalarOp 3 s v
        member _.Mul_M_V(A, x) = wrap A.Rows (ops.gemv (ctx, 0, A.Rows, A.Cols, h A.DataBuffer, h x, 0UL))
        member _.Mul_M_V_Add_V(A, x, y) = wrap A.Rows (ops.gemv (ctx, 0, A.Rows, A.Cols, h A.DataBuffer, h x, h y))
        member _.Mul_V_M(x, A) = wrap A.Cols (ops.gemv (ctx, 1, A.Rows, A.Cols, h A.DataBuffer, h x, 0UL))
        member _.Solve_M_V(A, b) = ops.gesv (ctx, A.Rows, h A.DataBuffer, h b) |> Option.map (wrap A.Rows)
        member _.SolveSymmetric_M_V(A, b) = ops.sysv (ctx, A.Rows, h A.DataBuffer, h b) |> Option.map (wrap A.Rows)
        member _.Diagonal_M A = wrap (min A.Rows A.Cols) (ops.diag (ctx, A.Rows, A.Cols, h A.DataBuffer))
        // only the op-code crosses; the closure argument is ignored and never executed
        member _.Map_F_V(op, _f, v) = wrap v.Length (ops.map (ctx, OpCode.ofMapOp op, h v))
        member _.Map_F_S_V(s, op, _f, v) = wrap v.Length (ops.map_s (ctx, OpCode.ofMapOp op, s, h v))
        member _.Map2_F_V_V(op, _f, a, b) = map2 (OpCode.ofMapOp op) a b
        member _.ReshapeCopy_MRows_V A = (A.DataBuffer.GetValues 0 (A.Rows * A.Cols)).DeepCopy()
        member _.Mul_Out_V_V(a, b) = view (wrap (a.Length * b.Length) (ops.ger (ctx, h a, h b))) [| int64 a.Length; int64 b.Length |]
        member _.Add_M_M(a, b) = view (map2 2 a.DataBuffer b.DataBuffer) (if a.Length > 0 then a.Shape else b.Shape)
        member _.Add_M_M_InPlace(a, b) =
            let d = dev a.DataBuffer
            ops.add_inplace (ctx, d.Handle, h b.DataBuffer); d.Invalidate(); a
        member _.Add_S_M(s, a) = view (scalarOp 0 s a.DataBuffer) a.Shape
        member _.Add_V_MCols(v, a) = view (wrap (a.Rows * a.Cols) (ops.add_v_mcols (ctx, a.Rows, a.Cols, h v, h a.DataBuffer))) a.Shape
        member _.Sub_M_M(a, b) = view (map2 3 a.DataBuffer b.DataBuffer) (if a.Length > 0 then a.Shape else b.Shape)
        member _.Sub_M_S(a, s) = view (scalarOp 2 s a.DataBuffer) a.Shape
        member _.Sub_S_M(s, a) = view (scalarOp 1 s a.DataBuffer) a.Shape
        member _.Mul_M_M(a, b) =
            if a.Cols <> b.Rows then ErrorMessages.InvalidArgMColsMRows()
            view (wrap (a.Rows * b.Cols) (ops.gemm (ctx, 0, 0, a.Rows, b.Cols, a.Cols, h a.DataBuffer, h b.DataBuffer, 0UL))) [| int64 a.Rows; int64 b.Cols |]
        member _.Mul_S_M(s, a) = view (scalarOp 3 s a.DataBuffer) a.Shape
        member _.Mul_M_M_Add_V_MCols(a, b, v) =
            view (wrap (a.Rows * b.Cols) (ops.gemm (ctx, 0, 0, a.Rows, b.Cols, a.Cols, h a.DataBuffer, h b.DataBuffer, h v))) [| int64 a.Rows; int64 b.Cols |]
        member _.Add_M_Colwise_V_InPlace(a, v) =
            let d = dev v
            ops.colsum_inplace (ctx, a.Rows, a.Cols, h a.DataBuffer, d.Handle); d.Invalidate(); v
        member _.Mul_Had_M_M(a, b) = view (map2 0 a.DataBuffer b.DataBuffer) a.Shape
        member _.Inverse_M a = ops.getri (ctx, a.Rows, h a.DataBuffer) |> Option.map (fun r -> view (wrap (a.Rows * a.Rows) r) a.Shape)
        member _.Det_M a = ops.det (ctx, a.Rows, h a.DataBuffer)
        member _.Transpose_M a = view (wrap (a.Rows * a.Cols) (ops.transpose (ctx, a.Rows, a.Cols, h a.DataBuffer))) [| int64 a.Cols; int64 a.Rows |]
        member _.Permute_M(a, dims) =
            view (wrap (Array.fold (*) 1L a.Shape |> int) (ops.permute (ctx, a.Shape, dims, h a.DataBuffer))) (dims |> Array.map (fun d -> a.Shape.[d]))
        member _.Reshape_M(a, shape) = ShapedDataBufferView<'T>(a.DataBuffer, shape)
        member b.Map_F_M(op, f, a) = view ((b :> Backend<'T>).Map_F_V(op, f, a.DataBuffer)) a.Shape
        member b.Map_F_S_M(s, op, f, a) = view ((b :> Backend<'T>).Map_F_S_V(s, op, f, a.DataBuffer)) a.Shape
        member _.Map2_F_M_M(op, _f, a, b) = view (map2 (OpCode.ofMapOp op) a.DataBuffer b.DataBuffer) a.Shape
        member _.ReshapeCopy_V_MRows(m, v) = view (v.DeepCopy()) [| int64 m; int64 (v.Length / m) |]
        member _.RepeatReshapeCopy_V_MRows(m, v) = view (wrap (m * v.Length) (ops.repeat_rows (ctx, m, h v))) [| int64 m; int64 v.Length |]
        member _.RepeatReshapeCopy_V_MCols(n, v) = view (wrap (n * v.Length) (ops.repeat_cols (ctx, n, h v))) [| int64 v.Length; int64 n |]
        // consumer hook (Backend.fs:137-139): the opaque payload is the consumer's own ICudaCustomOp, handed the backend
        member b.CustomOp_DM_Forward(a, info) =
            match info with
            | :? ICudaCustomOp<'T> as op -> op.Forward(b :> Backend<'T>, a)
            | _ -> invalidArg "customInfo" "expected an ICudaCustomOp<'T>"
        member b.CustomOp_DM_Backward(origin, adjoint, primal, info) =
            match info with
            | :? ICudaCustomOp<'T> as op -> op.Backward(b :> Backend<'T>, origin, adjoint, primal)
            | _ -> invalidArg "customInfo" "expected an ICudaCustomOp<'T>"

/// IBackendProvider (Config.fs:53-54): one context per process / GPU; keyed on 'T only, because
/// call sites pass buffers, views, boxed scalars and AD values alike (AD.Float32.fs:70).
type CudaBackendProvider(device : int, flags : uint32) =
    let mutable ctx = 0n
    do Native.check (Native.dsc_create(device, flags, &ctx))
    let f32 = { BackendHandle = CudaBackend<float32>(ctx, Ops.float32Ops, 0.f) :> Backend<float32>
                Epsilon = 0.00001f; EpsilonRec = 100000.f; EpsilonRec2 = 50000.f
                FixedPointEpsilon = 0.01f; FixedPointMaxIterations = 100; VisualisationContrast = 1.2f }
    let f64 = { BackendHandle = CudaBackend<float>(ctx, Ops.float64Ops, 0.) :> Backend<float>
                Epsilon = 0.00001; EpsilonRec = 100000.; EpsilonRec2 = 50000.
                FixedPointEpsilon = 0.01; FixedPointMaxIterations = 100; VisualisationContrast = 1.2 }
    new(device) = CudaBackendProvider(device, 0u)
    member _.Context = ctx
    member _.Synchronise() = Native.check (Native.dsc_sync ctx)
    interface IBackendProvider with
        member _.GetBackend<'T>(_value : obj) : BackendConfig<'T> =
            if typeof<'T> = typeof<float32> then (box f32) :?> BackendConfig<'T>
            elif typeof<'T> = typeof<float> then (box f64) :?> BackendConfig<'T>
            else invalidArg "" "Unsupported type, try float32 or float"   // Config.fs:120
// usage:  GlobalConfig.BackendProvider <- CudaBackendProvider(0)      (Config.fs:89-91)
