// Backend.Cuda.fs — managed shim: Backend<'T>, ISigmaDiffDataBuffer<'T> and IBackendProvider over
// libdiffsharp_cuda.so (P/Invoke), to be compiled into the CONSUMER of Sigma.DiffSharp next to (or
// instead of) its OpenBLAS backend.
//
// STATUS: NEVER COMPILED OR RUN.  This image has no .NET / Mono / F# toolchain.  The file documents the
// binding a maintainer would add; its twin that IS executed by the test-suite is the Python ctypes
// layer (sigma/diffsharp_b200/_lib.py, buffers.py, backend.py), which makes the same calls in the
// same order.  float32 shown; the float64 twin replaces `dsc_s*` by `dsc_d*` and float32 by float.
namespace DiffSharp.Backend.Cuda

open System
open System.Runtime.InteropServices
open System.Runtime.CompilerServices
open DiffSharp.Util
open DiffSharp.Backend
open DiffSharp.Config

module Native =
    [<Literal>]
    let Lib = "diffsharp_cuda" // libdiffsharp_cuda.so on the library path

    [<DllImport(Lib)>] extern int dsc_create(int device, uint32 flags, nativeint& ctx)
    [<DllImport(Lib)>] extern int dsc_destroy(nativeint ctx)
    [<DllImport(Lib)>] extern int dsc_sync(nativeint ctx)
    [<DllImport(Lib)>] extern nativeint dsc_last_error()
    [<DllImport(Lib)>] extern int dsc_buf_alloc(nativeint ctx, int dtype, int n, uint64& h)
    [<DllImport(Lib)>] extern int dsc_buf_alloc_fill(nativeint ctx, int dtype, int n, double v, uint64& h)
    [<DllImport(Lib)>] extern int dsc_buf_free(uint64 h)
    [<DllImport(Lib)>] extern int dsc_buf_upload(uint64 h, int off, float32[] host, int n)
    [<DllImport(Lib)>] extern int dsc_buf_download(uint64 h, int off, float32[] host, int n)
    [<DllImport(Lib)>] extern int dsc_buf_view(uint64 h, int off, int n, uint64& out)
    [<DllImport(Lib)>] extern int dsc_buf_copy(uint64 h, uint64& out)
    [<DllImport(Lib)>] extern int dsc_buf_gather2d(uint64 h, int rows, int cols, int r0, int r1, int c0, int c1, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sgemm(nativeint ctx, int ta, int tb, int m, int n, int k, uint64 A, uint64 B, uint64 bias, uint64& C)
    [<DllImport(Lib)>] extern int dsc_sgemv(nativeint ctx, int trans, int m, int n, uint64 A, uint64 x, uint64 y, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sger(nativeint ctx, uint64 x, uint64 y, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sdot(nativeint ctx, uint64 x, uint64 y, float32& out)
    [<DllImport(Lib)>] extern int dsc_sasum(nativeint ctx, uint64 x, float32& out)
    [<DllImport(Lib)>] extern int dsc_snrm2(nativeint ctx, uint64 x, float32& out)
    [<DllImport(Lib)>] extern int dsc_samax(nativeint ctx, uint64 x, float32& out)
    [<DllImport(Lib)>] extern int dsc_ssum(nativeint ctx, uint64 x, float32& out)
    [<DllImport(Lib)>] extern int dsc_sargmax(nativeint ctx, uint64 x, int& out)
    [<DllImport(Lib)>] extern int dsc_sargmin(nativeint ctx, uint64 x, int& out)
    [<DllImport(Lib)>] extern int dsc_smap(nativeint ctx, int op, uint64 x, uint64& y)
    [<DllImport(Lib)>] extern int dsc_smap_s(nativeint ctx, int op, float32 s, uint64 x, uint64& y)
    [<DllImport(Lib)>] extern int dsc_smap2(nativeint ctx, int op, uint64 a, uint64 b, uint64& y)
    [<DllImport(Lib)>] extern int dsc_sscalar_op(nativeint ctx, int kind, float32 s, uint64 x, uint64& y)
    [<DllImport(Lib)>] extern int dsc_sadd_inplace(nativeint ctx, uint64 dst, uint64 src)
    [<DllImport(Lib)>] extern int dsc_sadd_inplace_range(nativeint ctx, uint64 src, int soff, uint64 dst, int doff, int n)
    [<DllImport(Lib)>] extern int dsc_sfused_bwd(nativeint ctx, int kind, uint64 dA, uint64 p, uint64& out)
    [<DllImport(Lib)>] extern int dsc_stranspose(nativeint ctx, int m, int n, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_spermute(nativeint ctx, int rank, int64[] shape, int[] perm, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_srepeat_rows(nativeint ctx, int m, uint64 v, uint64& out)
    [<DllImport(Lib)>] extern int dsc_srepeat_cols(nativeint ctx, int n, uint64 v, uint64& out)
    [<DllImport(Lib)>] extern int dsc_scolsum_inplace(nativeint ctx, int m, int n, uint64 M, uint64 v)
    [<DllImport(Lib)>] extern int dsc_sadd_v_mcols(nativeint ctx, int m, int n, uint64 v, uint64 M, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sdiag(nativeint ctx, int m, int n, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sgesv(nativeint ctx, int n, uint64 A, uint64 b, uint64& x)
    [<DllImport(Lib)>] extern int dsc_ssysv(nativeint ctx, int n, uint64 A, uint64 b, uint64& x)
    [<DllImport(Lib)>] extern int dsc_sgetri(nativeint ctx, int n, uint64 A, uint64& out)
    [<DllImport(Lib)>] extern int dsc_sdet(nativeint ctx, int n, uint64 A, float32& out)

    let DSC_E_SHAPE, DSC_E_UNSUPPORTED_OP, DSC_E_SINGULAR = -2, -6, -7
    /// status -> .NET exception, mirroring the reference's own error style (Util.fs:117-131)
    let check (status : int) =
        if status <> 0 then
            let msg = Marshal.PtrToStringAnsi(dsc_last_error())
            match status with
            | s when s = DSC_E_SHAPE -> invalidArg "" msg
            | s when s = DSC_E_UNSUPPORTED_OP -> raise (NotSupportedException msg) // closures are never executed
            | _ -> failwithf "libdiffsharp_cuda: %s" msg

/// Device-resident ISigmaDiffDataBuffer (Util.fs:133-141).  Owns one dsc_buf_t.
type CudaDataBuffer(ctx : nativeint, handle : uint64, length : int, offset : int) =
    let mutable host : float32[] = null // lazy mirror, only for .Data/.SubData (printing, ToArray)
    member _.Handle = handle
    member _.Ctx = ctx
    override _.Finalize() = Native.dsc_buf_free handle |> ignore
    interface ISigmaDiffDataBuffer<float32> with
        member _.Length = length
        member _.Offset = offset
        member d.SubData =
            let a = Array.zeroCreate length
            if length > 0 then Native.check (Native.dsc_buf_download(handle, 0, a, length))
            a
        member d.Data =
            let sub = (d :> ISigmaDiffDataBuffer<float32>).SubData
            let a = Array.zeroCreate (offset + length)
            Array.blit sub 0 a offset length
            a
        member _.GetValues startIndex len =
            let mutable h = 0UL
            Native.check (Native.dsc_buf_view(handle, startIndex, len, &h))
            CudaDataBuffer(ctx, h, len, offset + startIndex) :> ISigmaDiffDataBuffer<float32>
        member _.GetStackedValues rows cols r0 r1 c0 c1 =
            let mutable h = 0UL
            Native.check (Native.dsc_buf_gather2d(handle, rows, cols, r0, r1, c0, c1, &h))
            CudaDataBuffer(ctx, h, (r1 - r0 + 1) * (c1 - c0 + 1), 0) :> ISigmaDiffDataBuffer<float32>
        member _.DeepCopy() =
            let mutable h = 0UL
            Native.check (Native.dsc_buf_copy(handle, &h))
            CudaDataBuffer(ctx, h, length, 0) :> ISigmaDiffDataBuffer<float32>
        member d.ShallowCopy() = (d :> ISigmaDiffDataBuffer<float32>).GetValues 0 length

/// Backend<float32> over the C ABI.  One P/Invoke per member.
type CudaBackendFloat32(ctx : nativeint) =
    // CreateZeroArray / CreateValueArray return managed arrays registered here as "n copies of v" so that
    // CreateDataBuffer can fill on the device instead of touching the host pages (SURVEY §7.3).
    let constants = ConditionalWeakTable<float32[], obj>()
    let dev (b : ISigmaDiffDataBuffer<float32>) : CudaDataBuffer =
        match b with
        | :? CudaDataBuffer as c -> c
        | _ -> // foreign buffer (DVector.ZeroN / DNDArray.Zero, AD.Float32.fs:706-707,1833): upload on first touch
            let mutable h = 0UL
            Native.check (Native.dsc_buf_alloc(ctx, 0, b.Length, &h))
            if b.Length > 0 then Native.check (Native.dsc_buf_upload(h, 0, b.SubData, b.Length))
            CudaDataBuffer(ctx, h, b.Length, 0)
    let wrap n (h : uint64) = CudaDataBuffer(ctx, h, n, 0) :> ISigmaDiffDataBuffer<float32>
    let view (b : ISigmaDiffDataBuffer<float32>) ([<ParamArray>] shape : int64[]) = ShapedDataBufferView<float32>(b, shape)
    let map2 op (a : ISigmaDiffDataBuffer<float32>) (b : ISigmaDiffDataBuffer<float32>) =
        let mutable h = 0UL
        Native.check (Native.dsc_smap2(ctx, op, (dev a).Handle, (dev b).Handle, &h))
        wrap (max a.Length b.Length) h
    let scalarOp kind s (v : ISigmaDiffDataBuffer<float32>) =
        let mutable h = 0UL
        Native.check (Native.dsc_sscalar_op(ctx, kind, s, (dev v).Handle, &h))
        wrap v.Length h
    interface Backend<float32> with
        member _.CreateDataBuffer a =
            match constants.TryGetValue a with
            | true, (:? float32 as v) ->
                let mutable h = 0UL
                Native.check (Native.dsc_buf_alloc_fill(ctx, 0, a.Length, float v, &h))
                wrap a.Length h
            | _ ->
                let mutable h = 0UL
                Native.check (Native.dsc_buf_alloc(ctx, 0, a.Length, &h))
                if a.Length > 0 then Native.check (Native.dsc_buf_upload(h, 0, a, a.Length))
                wrap a.Length h
        member _.CreateUninitialisedArray n = Array.zeroCreate n
        member _.CreateZeroArray n = let a = Array.zeroCreate n in constants.Add(a, box 0.f); a
        member _.CreateValueArray(n, v) = let a = Array.create n v in constants.Add(a, box v); a
        member _.Mul_Dot_V_V(a, b) = let mutable r = 0.f in Native.check (Native.dsc_sdot(ctx, (dev a).Handle, (dev b).Handle, &r)); r
        member _.L1Norm_V a = let mutable r = 0.f in Native.check (Native.dsc_sasum(ctx, (dev a).Handle, &r)); r
        member _.L2Norm_V a = let mutable r = 0.f in Native.check (Native.dsc_snrm2(ctx, (dev a).Handle, &r)); r
        member _.SupNorm_V a = let mutable r = 0.f in Native.check (Native.dsc_samax(ctx, (dev a).Handle, &r)); r
        member _.Sum_V a = let mutable r = 0.f in Native.check (Native.dsc_ssum(ctx, (dev a).Handle, &r)); r
        member _.Sum_M a = let mutable r = 0.f in Native.check (Native.dsc_ssum(ctx, (dev a).Handle, &r)); r
        member _.MaxIndex_V a = let mutable r = 0 in Native.check (Native.dsc_sargmax(ctx, (dev a).Handle, &r)); r
        member _.MinIndex_V a = let mutable r = 0 in Native.check (Native.dsc_sargmin(ctx, (dev a).Handle, &r)); r
        member _.Add_V_V(a, b) = map2 2 a b
        member _.Sub_V_V(a, b) = map2 3 a b
        member _.Add_V_V_InPlace(src, so, dst, dof, n) =
            Native.check (Native.dsc_sadd_inplace_range(ctx, (dev src).Handle, so, (dev dst).Handle, dof, n)); dst
        member _.Add_S_V(s, v) = scalarOp 0 s v
        member _.Sub_S_V(s, v) = scalarOp 1 s v
        member _.Sub_V_S(v, s) = scalarOp 2 s v
        member _.Mul_S_V(s, v) = scalarOp 3 s v
        member _.Mul_M_V(A, x) =
            let mutable h = 0UL
            Native.check (Native.dsc_sgemv(ctx, 0, A.Rows, A.Cols, (dev A.DataBuffer).Handle, (dev x).Handle, 0UL, &h)); wrap A.Rows h
        member _.Mul_M_V_Add_V(A, x, y) =
            let mutable h = 0UL
            Native.check (Native.dsc_sgemv(ctx, 0, A.Rows, A.Cols, (dev A.DataBuffer).Handle, (dev x).Handle, (dev y).Handle, &h)); wrap A.Rows h
        member _.Mul_V_M(x, A) =
            let mutable h = 0UL
            Native.check (Native.dsc_sgemv(ctx, 1, A.Rows, A.Cols, (dev A.DataBuffer).Handle, (dev x).Handle, 0UL, &h)); wrap A.Cols h
        member _.Solve_M_V(A, b) =
            let mutable h = 0UL
            match Native.dsc_sgesv(ctx, A.Rows, (dev A.DataBuffer).Handle, (dev b).Handle, &h) with
            | s when s = Native.DSC_E_SINGULAR -> None
            | s -> Native.check s; Some(wrap A.Rows h)
        member _.SolveSymmetric_M_V(A, b) =
            let mutable h = 0UL
            match Native.dsc_ssysv(ctx, A.Rows, (dev A.DataBuffer).Handle, (dev b).Handle, &h) with
            | s when s = Native.DSC_E_SINGULAR -> None
            | s -> Native.check s; Some(wrap A.Rows h)
        member _.Diagonal_M A =
            let mutable h = 0UL
            Native.check (Native.dsc_sdiag(ctx, A.Rows, A.Cols, (dev A.DataBuffer).Handle, &h)); wrap (min A.Rows A.Cols) h
        // only the op-code crosses; the closure argument is ignored and never executed
        member _.Map_F_V(op, _f, v) = let mutable h = 0UL in Native.check (Native.dsc_smap(ctx, int op, (dev v).Handle, &h)); wrap v.Length h
        member _.Map_F_S_V(s, op, _f, v) = let mutable h = 0UL in Native.check (Native.dsc_smap_s(ctx, int op, s, (dev v).Handle, &h)); wrap v.Length h
        member _.Map2_F_V_V(op, _f, a, b) = map2 (int op) a b
        member _.ReshapeCopy_MRows_V A = (A.DataBuffer.GetValues 0 (A.Rows * A.Cols)).DeepCopy()
        member _.Mul_Out_V_V(a, b) =
            let mutable h = 0UL
            Native.check (Native.dsc_sger(ctx, (dev a).Handle, (dev b).Handle, &h)); view (wrap (a.Length * b.Length) h) [| int64 a.Length; int64 b.Length |]
        member _.Add_M_M(a, b) = view (map2 2 a.DataBuffer b.DataBuffer) (if a.Length > 0 then a.Shape else b.Shape)
        member _.Add_M_M_InPlace(a, b) = Native.check (Native.dsc_sadd_inplace(ctx, (dev a.DataBuffer).Handle, (dev b.DataBuffer).Handle)); a
        member _.Add_S_M(s, a) = view (scalarOp 0 s a.DataBuffer) a.Shape
        member _.Add_V_MCols(v, a) =
            let mutable h = 0UL
            Native.check (Native.dsc_sadd_v_mcols(ctx, a.Rows, a.Cols, (dev v).Handle, (dev a.DataBuffer).Handle, &h)); view (wrap (a.Rows * a.Cols) h) a.Shape
        member _.Sub_M_M(a, b) = view (map2 3 a.DataBuffer b.DataBuffer) (if a.Length > 0 then a.Shape else b.Shape)
        member _.Sub_M_S(a, s) = view (scalarOp 2 s a.DataBuffer) a.Shape
        member _.Sub_S_M(s, a) = view (scalarOp 1 s a.DataBuffer) a.Shape
        member _.Mul_M_M(a, b) =
            if a.Cols <> b.Rows then ErrorMessages.InvalidArgMColsMRows()
            let mutable h = 0UL
            Native.check (Native.dsc_sgemm(ctx, 0, 0, a.Rows, b.Cols, a.Cols, (dev a.DataBuffer).Handle, (dev b.DataBuffer).Handle, 0UL, &h))
            view (wrap (a.Rows * b.Cols) h) [| int64 a.Rows; int64 b.Cols |]
        member _.Mul_S_M(s, a) = view (scalarOp 3 s a.DataBuffer) a.Shape
        member _.Mul_M_M_Add_V_MCols(a, b, v) =
            let mutable h = 0UL
            Native.check (Native.dsc_sgemm(ctx, 0, 0, a.Rows, b.Cols, a.Cols, (dev a.DataBuffer).Handle, (dev b.DataBuffer).Handle, (dev v).Handle, &h))
            view (wrap (a.Rows * b.Cols) h) [| int64 a.Rows; int64 b.Cols |]
        member _.Add_M_Colwise_V_InPlace(a, v) =
            Native.check (Native.dsc_scolsum_inplace(ctx, a.Rows, a.Cols, (dev a.DataBuffer).Handle, (dev v).Handle)); v
        member _.Mul_Had_M_M(a, b) = view (map2 0 a.DataBuffer b.DataBuffer) a.Shape
        member _.Inverse_M a =
            let mutable h = 0UL
            match Native.dsc_sgetri(ctx, a.Rows, (dev a.DataBuffer).Handle, &h) with
            | s when s = Native.DSC_E_SINGULAR -> None
            | s -> Native.check s; Some(view (wrap (a.Rows * a.Rows) h) a.Shape)
        member _.Det_M a =
            let mutable r = 0.f
            match Native.dsc_sdet(ctx, a.Rows, (dev a.DataBuffer).Handle, &r) with
            | s when s = Native.DSC_E_SINGULAR -> None
            | s -> Native.check s; Some r
        member _.Transpose_M a =
            let mutable h = 0UL
            Native.check (Native.dsc_stranspose(ctx, a.Rows, a.Cols, (dev a.DataBuffer).Handle, &h))
            view (wrap (a.Rows * a.Cols) h) [| int64 a.Cols; int64 a.Rows |]
        member _.Permute_M(a, dims) =
            let mutable h = 0UL
            Native.check (Native.dsc_spermute(ctx, a.Shape.Length, a.Shape, dims, (dev a.DataBuffer).Handle, &h))
            view (wrap (Array.fold (*) 1L a.Shape |> int) h) (dims |> Array.map (fun d -> a.Shape.[d]))
        member _.Reshape_M(a, shape) = ShapedDataBufferView<float32>(a.DataBuffer, shape)
        member b.Map_F_M(op, f, a) = view ((b :> Backend<float32>).Map_F_V(op, f, a.DataBuffer)) a.Shape
        member b.Map_F_S_M(s, op, f, a) = view ((b :> Backend<float32>).Map_F_S_V(s, op, f, a.DataBuffer)) a.Shape
        member _.Map2_F_M_M(op, _f, a, b) = view (map2 (int op) a.DataBuffer b.DataBuffer) a.Shape
        member _.ReshapeCopy_V_MRows(m, v) = view (v.DeepCopy()) [| int64 m; int64 (v.Length / m) |]
        member _.RepeatReshapeCopy_V_MRows(m, v) =
            let mutable h = 0UL
            Native.check (Native.dsc_srepeat_rows(ctx, m, (dev v).Handle, &h)); view (wrap (m * v.Length) h) [| int64 m; int64 v.Length |]
        member _.RepeatReshapeCopy_V_MCols(n, v) =
            let mutable h = 0UL
            Native.check (Native.dsc_srepeat_cols(ctx, n, (dev v).Handle, &h)); view (wrap (n * v.Length) h) [| int64 v.Length; int64 n |]
        member _.CustomOp_DM_Forward(a, _info) = a       // consumer hook: supplied by the consumer
        member _.CustomOp_DM_Backward(_o, adj, _p, _info) = adj

/// IBackendProvider (Config.fs:53-54): one context per process / GPU; keyed on 'T only, because
/// call sites pass buffers, views, boxed scalars and AD values alike (AD.Float32.fs:70).
type CudaBackendProvider(device : int) =
    let mutable ctx = 0n
    do Native.check (Native.dsc_create(device, 0u, &ctx))
    let f32 = { BackendHandle = CudaBackendFloat32(ctx) :> Backend<float32>
                Epsilon = 0.00001f; EpsilonRec = 100000.f; EpsilonRec2 = 50000.f
                FixedPointEpsilon = 0.01f; FixedPointMaxIterations = 100; VisualisationContrast = 1.2f }
    interface IBackendProvider with
        member _.GetBackend<'T>(_value : obj) : BackendConfig<'T> =
            if typeof<'T> = typeof<float32> then (box f32) :?> BackendConfig<'T>
            else invalidArg "" "Unsupported type, try float32 or float"
// usage:  GlobalConfig.BackendProvider <- CudaBackendProvider(0)      (Config.fs:89-91)
