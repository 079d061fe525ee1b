// SoftbodyB200.cs -- managed side of the drop-in boundary (NOT compiled in this repository: the
// build image has no C#/.NET/Unity toolchain; reviewed by eye against include/softbody_b200.h and
// guarded at run time by sb_abi_check).
//
// The upstream MonoBehaviour this replaces is NOT IN MOUNT (/root/reference/README.md:1 is the whole
// reference).  BASELINE.json:5 names its surface -- a Step call and the inspector fields stiffness,
// damping, substeps, iterations -- and that is what this component keeps.  Everything between
// Start() and OnDestroy() is a thin P/Invoke into libsoftbody_b200.so (Assets/Plugins/x86_64/).
using System;
using System.Runtime.InteropServices;
using UnityEngine;

[StructLayout(LayoutKind.Sequential)]
public struct SbParams            // sb_params, 48 bytes
{
    public float dt;
    public int substeps;
    public int iterations;
    public float stiffnessDistance;
    public float stiffnessVolume;
    public float damping;
    public float friction;
    public float gravityX, gravityY, gravityZ;
    public float groundY;
    public int flags;
}

[StructLayout(LayoutKind.Sequential)]
public struct SbMeshDesc          // sb_mesh_desc, 112 bytes
{
    public IntPtr posXyz, tets, surfTris, invMass, stream, edges;
    public uint nVerts, nTets, nTris;
    public float density;
    public int device, tileCap, maxTilePasses, blockThreads, laterTileCap, hostThreads, roundWidth, attachEdges, tilings, nGhostVerts;
    public uint nEdges;
    public int distRanks;
}

[StructLayout(LayoutKind.Sequential)]
public struct SbCollider          // sb_collider, 48 bytes: kind 0 sphere / 1 capsule / 2 box
{
    public int kind;
    public float friction;
    public float p0, p1, p2, p3, p4, p5, p6, p7, p8, p9;

    public static SbCollider From(Collider c, float friction)   // world-space pose of a Unity collider
    {
        var t = c.transform;
        var k = new SbCollider { friction = friction };
        if (c is SphereCollider s)
        {
            var o = t.TransformPoint(s.center);
            float r = s.radius * Mathf.Max(Mathf.Abs(t.lossyScale.x), Mathf.Abs(t.lossyScale.y), Mathf.Abs(t.lossyScale.z));
            k.kind = 0; k.p0 = o.x; k.p1 = o.y; k.p2 = o.z; k.p3 = r;
        }
        else if (c is CapsuleCollider cap)
        {
            var axis = cap.direction == 0 ? Vector3.right : cap.direction == 1 ? Vector3.up : Vector3.forward;
            float half = Mathf.Max(0f, cap.height * 0.5f - cap.radius);
            var a = t.TransformPoint(cap.center - axis * half);
            var b = t.TransformPoint(cap.center + axis * half);
            k.kind = 1; k.p0 = a.x; k.p1 = a.y; k.p2 = a.z; k.p3 = cap.radius * Mathf.Abs(t.lossyScale.x);
            k.p4 = b.x; k.p5 = b.y; k.p6 = b.z;
        }
        else if (c is BoxCollider box)
        {
            var o = t.TransformPoint(box.center);
            var h = Vector3.Scale(box.size * 0.5f, t.lossyScale);
            var q = t.rotation;
            k.kind = 2; k.p0 = o.x; k.p1 = o.y; k.p2 = o.z; k.p3 = Mathf.Abs(h.x); k.p4 = Mathf.Abs(h.y); k.p5 = Mathf.Abs(h.z);
            k.p6 = q.x; k.p7 = q.y; k.p8 = q.z; k.p9 = q.w;
        }
        else throw new NotSupportedException("sphere, capsule and box colliders only");
        return k;
    }
}

internal static class SbNative
{
    const string Lib = "softbody_b200";
    [DllImport(Lib)] public static extern int sb_abi_check(out uint version, out uint sizeofParams, out uint sizeofDesc, out uint sizeofInfo);
    [DllImport(Lib)] public static extern void sb_default_params(out SbParams p);
    [DllImport(Lib)] public static extern int sb_create(ref SbMeshDesc mesh, ref SbParams prm, out IntPtr handle);
    [DllImport(Lib)] public static extern int sb_destroy(IntPtr h);
    [DllImport(Lib)] public static extern int sb_set_params(IntPtr h, ref SbParams prm);
    [DllImport(Lib)] public static extern int sb_set_colliders(IntPtr h, float[] spheresXyzr, uint n);
    [DllImport(Lib)] public static extern int sb_set_colliders_ex(IntPtr h, [In] SbCollider[] colliders, uint n);
    [DllImport(Lib)] public static extern int sb_step(IntPtr h, float dt);
    // render mesh embedded in the tets, snapshots, mesh ingest (all optional for the hot path)
    [DllImport(Lib)] public static extern int sb_skin_bind(IntPtr h, IntPtr renderPosXyz, uint nRenderVerts, int[] renderTris, uint nRenderTris);
    [DllImport(Lib)] public static extern int sb_read_skinned(IntPtr h, IntPtr dstPosXyz, IntPtr dstNrmXyz, uint nRenderVerts);
    [DllImport(Lib)] public static extern int sb_save_state(IntPtr h, string path);
    [DllImport(Lib)] public static extern int sb_load_state(IntPtr h, string path, int applyParams);
    [DllImport(Lib)] public static extern int sb_tetmesh_from_surface(IntPtr surfPosXyz, uint nVerts, int[] surfTris, uint nTris, float spacing, out IntPtr mesh);
    [DllImport(Lib)] public static extern int sb_tetmesh_snap_to_surface(IntPtr mesh, IntPtr surfPosXyz, uint nVerts, int[] surfTris, uint nTris, float maxDist, out uint nMoved);
    [DllImport(Lib)] public static extern int sb_tetmesh_load(string path, out IntPtr mesh);
    [DllImport(Lib)] public static extern int sb_tetmesh_desc(IntPtr mesh, out SbMeshDesc desc);
    [DllImport(Lib)] public static extern int sb_tetmesh_free(IntPtr mesh);
    [DllImport(Lib)] public static extern IntPtr sb_ingest_last_error();
    [DllImport(Lib)] public static extern int sb_read_positions(IntPtr h, IntPtr dstXyz, uint nVerts);
    [DllImport(Lib)] public static extern int sb_read_normals(IntPtr h, IntPtr dstXyz, uint nVerts);
    [DllImport(Lib)] public static extern IntPtr sb_last_error(IntPtr h);
}

public class SoftbodyB200 : MonoBehaviour
{
    // inspector fields (names per BASELINE.json:5; units and defaults are this repo's, see DESIGN.md)
    public float stiffness = float.PositiveInfinity;        // edge springs, N/m; +inf = rigid
    public float volumeStiffness = float.PositiveInfinity;
    public float damping = 0f;                               // 1/s
    public int substeps = 10;
    public int iterations = 10;
    public float friction = 0f;
    public float density = 1000f;
    public int device = 0;

    // tetrahedral mesh of the body (rest pose) and its render surface
    public Vector3[] restPositions;
    public int[] tets;           // 4 per tet
    public int[] surfaceTriangles;

    IntPtr handle = IntPtr.Zero;
    SbParams prm;
    Vector3[] positions, normals;
    GCHandle posPin, nrmPin;
    Mesh mesh;

    static void Check(int rc, IntPtr h)
    {
        if (rc != 0) throw new InvalidOperationException("softbody_b200 error " + rc + ": " + Marshal.PtrToStringAnsi(SbNative.sb_last_error(h)));
    }

    void FillParams()
    {
        prm.dt = Time.fixedDeltaTime;
        prm.substeps = substeps; prm.iterations = iterations;
        prm.stiffnessDistance = stiffness; prm.stiffnessVolume = volumeStiffness;
        prm.damping = damping; prm.friction = friction;
        prm.gravityX = Physics.gravity.x; prm.gravityY = Physics.gravity.y; prm.gravityZ = Physics.gravity.z;
        prm.groundY = 0f;
    }

    void Start()
    {
        uint ver, sp, sd, si;
        SbNative.sb_abi_check(out ver, out sp, out sd, out si);
        if (ver != 3 || sp != Marshal.SizeOf(typeof(SbParams)) || sd != Marshal.SizeOf(typeof(SbMeshDesc)))
            throw new InvalidOperationException("softbody_b200 ABI mismatch");
        SbNative.sb_default_params(out prm);
        FillParams();
        var p = GCHandle.Alloc(restPositions, GCHandleType.Pinned);   // Vector3 == 3 packed floats
        var t = GCHandle.Alloc(tets, GCHandleType.Pinned);
        var s = GCHandle.Alloc(surfaceTriangles, GCHandleType.Pinned);
        try
        {
            var d = new SbMeshDesc {
                posXyz = p.AddrOfPinnedObject(), tets = t.AddrOfPinnedObject(), surfTris = s.AddrOfPinnedObject(),
                invMass = IntPtr.Zero, stream = IntPtr.Zero, edges = IntPtr.Zero,
                nVerts = (uint)restPositions.Length, nTets = (uint)(tets.Length / 4), nTris = (uint)(surfaceTriangles.Length / 3),
                density = density, device = device, maxTilePasses = -1 };
            Check(SbNative.sb_create(ref d, ref prm, out handle), IntPtr.Zero);
        }
        finally { p.Free(); t.Free(); s.Free(); }    // the library keeps no host pointer after sb_create
        positions = new Vector3[restPositions.Length];
        normals = new Vector3[restPositions.Length];
        posPin = GCHandle.Alloc(positions, GCHandleType.Pinned);
        nrmPin = GCHandle.Alloc(normals, GCHandleType.Pinned);
        mesh = GetComponent<MeshFilter>().mesh;
        mesh.MarkDynamic();
    }

    // Scene colliders the body should respond to (sphere / capsule / box), re-sent whenever they move.
    public Collider[] sceneColliders;
    public float colliderFriction = 0f;
    public void PushColliders()
    {
        int n = sceneColliders == null ? 0 : Mathf.Min(sceneColliders.Length, 16);
        var list = new SbCollider[Mathf.Max(n, 1)];
        for (int i = 0; i < n; i++) list[i] = SbCollider.From(sceneColliders[i], colliderFriction);
        Check(SbNative.sb_set_colliders_ex(handle, list, (uint)n), handle);
    }

    // the hot path: one native call per fixed tick
    public void Step(float dt) { Check(SbNative.sb_step(handle, dt), handle); }

    void FixedUpdate()
    {
        if (prm.substeps != substeps || prm.iterations != iterations || prm.stiffnessDistance != stiffness ||
            prm.stiffnessVolume != volumeStiffness || prm.damping != damping || prm.friction != friction)
        {
            FillParams();
            Check(SbNative.sb_set_params(handle, ref prm), handle);
        }
        Step(Time.fixedDeltaTime);
    }

    // per-frame mesh write-back
    void LateUpdate()
    {
        Check(SbNative.sb_read_positions(handle, posPin.AddrOfPinnedObject(), (uint)positions.Length), handle);
        Check(SbNative.sb_read_normals(handle, nrmPin.AddrOfPinnedObject(), (uint)normals.Length), handle);
        mesh.vertices = positions;      // when the render mesh is the tet mesh's surface, indexed by vertex id
        mesh.normals = normals;
    }

    void OnDestroy()
    {
        if (handle != IntPtr.Zero) SbNative.sb_destroy(handle);
        handle = IntPtr.Zero;
        if (posPin.IsAllocated) posPin.Free();
        if (nrmPin.IsAllocated) nrmPin.Free();
    }
}
