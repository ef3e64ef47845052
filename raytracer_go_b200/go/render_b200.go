//go:build cgo

// render_b200.go — the cgo bridge that makes librt_b200.so a drop-in for Camera.Render.
//
// Drop this file into the reference's `internal/` directory (it must live in package `internal`
// because Camera, BVH, Sphere and the materials keep their fields unexported) and replace the call
// `camera.Render(worldTree, f)` (main.go:103,129,159,191,224,288) by `camera.RenderB200(worldTree, f)`.
//
// STATUS: written against include/rt_b200.h but NOT compiled — the build image has no Go toolchain
// (SURVEY.md F4).  The identical C ABI is exercised by the ctypes host mirror
// (raytracer_go_b200/api.py) and the C++ mirror (raytracer_go_b200/host/), which the tests run.
//
// What it does, in order:
//  1. walks the world (a *BVH built by NewBVH, bvh.go:142, or a *World, hittables.go:39) by type
//     switch and fills pointer-free C arrays: spheres, quads (Quad and Box, hittables.go:138-216),
//     materials, textures, images — materials and textures de-duplicated by pointer identity;
//  2. copies the derived camera fields computed by Camera.init (camera.go:128-166) into rt_camera;
//  3. makes ONE call, rt_render, which returns width*height RGB8 triplets (GetPixelColor + gamma +
//     quantise, camera.go:254-263, 212-214, on the GPU);
//  4. writes the PPM with the reference's own header code and chunk writer (camera.go:183-191,
//     237-252), so the bytes on disk are produced by the existing file path.
package internal

/*
#cgo CFLAGS: -I${SRCDIR}/../b200/include
#cgo LDFLAGS: -L${SRCDIR}/../b200/lib -lrt_b200 -Wl,-rpath,${SRCDIR}/../b200/lib
#include <stdlib.h>
#include "rt_b200.h"
*/
import "C"

import (
	"fmt"
	"image"
	"image/png"
	"io"
	"runtime"
	"strconv"
	"strings"
	"unsafe"
)

// B200Options has no counterpart in the reference: its RNGs are clock-seeded (camera.go:170).
type B200Options struct {
	Seed   uint64
	Device int
	// Devices, when it lists more than one CUDA ordinal, renders on all of them from this one
	// call (rt_render_multi).  TileSplit gives every device interleaved scanlines instead of a
	// share of the samples: the image is then bit-identical to the single-GPU one.
	Devices   []int
	TileSplit bool
	// Binary writes a P6 PPM (3 bytes per pixel) instead of the reference's P3 text — the TODO at
	// camera.go:196; a 3840x2160 frame is 24.9 MB instead of ~100 MB of text.
	Binary bool
	// PNG writes a PNG through the standard library's encoder instead (takes precedence over Binary).
	PNG bool
}

type b200Flat struct {
	spheres   []C.rt_sphere
	quads     []C.rt_quad
	// Object ID = first-seen position in the walk.  For a *World that is World.hittables order — the order
	// World.Hit (hittables.go:55-72) breaks exact ties in.  For a *BVH (what main.go passes) the walk is left-then-
	// right through NewBVH's random-axis tree (bvh.go:147), so IDs follow that run's leaf order: the reference's own
	// BVH.Hit also resolves an exact tie towards the left subtree (bvh.go:235), i.e. in the same order.
	sphereIDs []C.uint32_t
	quadIDs   []C.uint32_t
	materials []C.rt_material
	textures  []C.rt_texture
	images    []image.Image
	perlins   []C.rt_perlin
	matIndex  map[Material]uint32
	texIndex  map[Texture]uint32
	seen      map[Hittable]bool
	seenQuads map[quadKey]bool
}

// A Quad is stored by value, so identity is its geometry AND its material: two coincident quads with different
// materials are two hittables.  (Material is an interface holding a pointer or a comparable struct: valid map key.)
type quadKey struct {
	Q, u, v  Vec3
	material Material
}

func (f *b200Flat) texture(t Texture) (uint32, error) {
	if i, ok := f.texIndex[t]; ok {
		return i, nil
	}
	var rec C.rt_texture
	switch tt := t.(type) {
	case SolidColor: // materials.go:151-153 (stored by value in main.go)
		v := tt.albedo.GetColor()
		rec.kind = C.RT_TEX_SOLID
		rec.a = [3]C.float{C.float(v.X), C.float(v.Y), C.float(v.Z)}
	case *SolidColor:
		v := tt.albedo.GetColor()
		rec.kind = C.RT_TEX_SOLID
		rec.a = [3]C.float{C.float(v.X), C.float(v.Y), C.float(v.Z)}
	case *Checkered: // materials.go:121-125
		e, o := tt.even.GetColor(), tt.odd.GetColor()
		rec.kind = C.RT_TEX_CHECKER
		rec.scale = C.float(tt.scale)
		rec.a = [3]C.float{C.float(e.X), C.float(e.Y), C.float(e.Z)}
		rec.b = [3]C.float{C.float(o.X), C.float(o.Y), C.float(o.Z)}
	case *ImageTexture: // materials.go:165-167
		rec.kind = C.RT_TEX_IMAGE
		rec.image = C.uint32_t(len(f.images))
		b := tt.img.Bounds()
		// image.At outside Bounds() returns the zero colour of the image's model; materials.go:181-186
		// indexes one past the edge whenever u clamps to 1 or v to 0, so the device needs that colour.
		r, g, bl, _ := tt.img.At(b.Max.X, b.Min.Y).RGBA()
		s := float32(1.0 / 65535.0)
		rec.oob = [3]C.float{C.float(float32(r) * s), C.float(float32(g) * s), C.float(float32(bl) * s)}
		f.images = append(f.images, tt.img)
	case *NoiseTexture: // materials.go:280-283: copy Perlin's tables so the device evaluates the same field
		rec.kind = C.RT_TEX_NOISE
		rec.scale = C.float(tt.scale)
		rec.image = C.uint32_t(len(f.perlins))
		var pt C.rt_perlin
		for k := 0; k < 256; k++ {
			v := tt.perlin.randVec3[k]
			pt.vec[k] = [3]C.float{C.float(v.X), C.float(v.Y), C.float(v.Z)}
			pt.perm_x[k] = C.uint8_t(tt.perlin.permX[k])
			pt.perm_y[k] = C.uint8_t(tt.perlin.permY[k])
			pt.perm_z[k] = C.uint8_t(tt.perlin.permZ[k])
		}
		f.perlins = append(f.perlins, pt)
	default:
		return 0, fmt.Errorf("b200: texture %T is outside the accelerated path", t)
	}
	f.textures = append(f.textures, rec)
	f.texIndex[t] = uint32(len(f.textures) - 1)
	return f.texIndex[t], nil
}

func (f *b200Flat) material(m Material) (uint32, error) {
	if i, ok := f.matIndex[m]; ok {
		return i, nil
	}
	var rec C.rt_material
	switch mm := m.(type) {
	case *Lambertian: // materials.go:19-21
		t, err := f.texture(mm.albedo)
		if err != nil {
			return 0, err
		}
		rec.kind, rec.texture = C.RT_MAT_LAMBERTIAN, C.uint32_t(t)
	case *Metal: // materials.go:44-47
		a := mm.albedo.GetColor()
		rec.kind = C.RT_MAT_METAL
		rec.albedo = [3]C.float{C.float(a.X), C.float(a.Y), C.float(a.Z)}
		rec.fuzz = C.float(mm.fuzz)
	case *Dielectric: // materials.go:77-79
		rec.kind, rec.ior = C.RT_MAT_DIELECTRIC, C.float(mm.refractiveIndex)
	case DiffuseLight: // materials.go:297-299 (value receiver)
		t, err := f.texture(mm.emit)
		if err != nil {
			return 0, err
		}
		rec.kind, rec.texture = C.RT_MAT_DIFFUSE_LIGHT, C.uint32_t(t)
	case *DiffuseLight:
		t, err := f.texture(mm.emit)
		if err != nil {
			return 0, err
		}
		rec.kind, rec.texture = C.RT_MAT_DIFFUSE_LIGHT, C.uint32_t(t)
	default:
		return 0, fmt.Errorf("b200: material %T is outside the accelerated path", m)
	}
	f.materials = append(f.materials, rec)
	f.matIndex[m] = uint32(len(f.materials) - 1)
	return f.matIndex[m], nil
}

// walk visits a *BVH (bvh.go:132-136), a *World (hittables.go:39-42) or a primitive.  A one-element
// BVH node holds the same object as both children (bvh.go:162-165), hence the `seen` set.
func (f *b200Flat) walk(h Hittable) error {
	switch hh := h.(type) {
	case *BVH:
		if err := f.walk(hh.left); err != nil {
			return err
		}
		return f.walk(hh.right)
	case *World:
		for _, c := range hh.hittables {
			if err := f.walk(c); err != nil {
				return err
			}
		}
		return nil
	case *Sphere: // hittables.go:78-83
		if f.seen[h] {
			return nil
		}
		f.seen[h] = true
		m, err := f.material(hh.Material)
		if err != nil {
			return err
		}
		f.sphereIDs = append(f.sphereIDs, C.uint32_t(len(f.spheres)+len(f.quads)))
		f.spheres = append(f.spheres, C.rt_sphere{
			cx: C.float(hh.Center.X), cy: C.float(hh.Center.Y), cz: C.float(hh.Center.Z),
			r: C.float(hh.Radius), material: C.uint32_t(m)})
		return nil
	case Quad: // hittables.go:138-147 (stored by value: NewQuad returns Quad, main.go:152)
		// a value type cannot be de-duplicated by pointer; a Quad reaches the walk once unless it is
		// the only element of a one-element BVH node (bvh.go:162-165), where left == right
		k := quadKey{hh.Q, hh.u, hh.v, hh.material}
		if f.seenQuads[k] {
			return nil
		}
		f.seenQuads[k] = true
		m, err := f.material(hh.material)
		if err != nil {
			return err
		}
		v3 := func(v Vec3) [3]C.float { return [3]C.float{C.float(v.X), C.float(v.Y), C.float(v.Z)} }
		f.quadIDs = append(f.quadIDs, C.uint32_t(len(f.spheres)+len(f.quads)))
		f.quads = append(f.quads, C.rt_quad{q: v3(hh.Q), u: v3(hh.u), v: v3(hh.v), material: C.uint32_t(m)})
		return nil
	default:
		return fmt.Errorf("b200: hittable %T is outside the accelerated path", h)
	}
}

// cArray copies a Go slice into C memory (cgo: memory handed to C may not contain Go pointers, and
// rt_scene_desc itself holds pointers, so the arrays it points to must not be Go memory).
func cArray[T any](s []T) (unsafe.Pointer, func()) {
	if len(s) == 0 {
		return nil, func() {}
	}
	n := C.size_t(len(s)) * C.size_t(unsafe.Sizeof(s[0]))
	p := C.malloc(n)
	copy(unsafe.Slice((*T)(p), len(s)), s)
	return p, func() { C.free(p) }
}

// rt_last_error() is thread-local in the library, and a goroutine may migrate between OS threads from one cgo call
// to the next: callB200 pins the goroutine for the failing call AND the fetch of its message.
func callB200(call func() C.int) error {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if rc := call(); rc != C.RT_OK {
		return fmt.Errorf("b200: %s (rt_status %d)", C.GoString(C.rt_last_error()), int(rc))
	}
	return nil
}

// RenderB200 has the signature and the output of Render (camera.go:180).
// RenderB200P6 is RenderB200 with the binary PPM writer.
func (c *Camera) RenderB200P6(world Hittable, writer io.Writer, opt ...B200Options) error {
	o := B200Options{Seed: 0xC0FFEE}
	if len(opt) > 0 {
		o = opt[0]
	}
	o.Binary = true
	return c.RenderB200(world, writer, o)
}

func (c *Camera) RenderB200(world Hittable, writer io.Writer, opt ...B200Options) error {
	o := B200Options{Seed: 0xC0FFEE}
	if len(opt) > 0 {
		o = opt[0]
	}
	f := &b200Flat{matIndex: map[Material]uint32{}, texIndex: map[Texture]uint32{}, seen: map[Hittable]bool{},
		seenQuads: map[quadKey]bool{}}
	if err := f.walk(world); err != nil {
		return err
	}

	// images: img.At(i,j).RGBA() at 16 bits (materials.go:186-189), row-major
	cimgs := make([]C.rt_image, len(f.images))
	var frees []func()
	defer func() {
		for _, fr := range frees {
			fr()
		}
	}()
	for k, img := range f.images {
		b := img.Bounds()
		w, h := b.Dx(), b.Dy()
		px := make([]uint16, 3*w*h)
		for j := 0; j < h; j++ {
			for i := 0; i < w; i++ {
				r, g, bl, _ := img.At(b.Min.X+i, b.Min.Y+j).RGBA()
				q := 3 * (j*w + i)
				px[q], px[q+1], px[q+2] = uint16(r), uint16(g), uint16(bl)
			}
		}
		p, fr := cArray(px)
		frees = append(frees, fr)
		cimgs[k] = C.rt_image{w: C.int32_t(w), h: C.int32_t(h), rgb16: (*C.uint16_t)(p)}
	}
	ps, fr := cArray(f.spheres)
	frees = append(frees, fr)
	pm, fr := cArray(f.materials)
	frees = append(frees, fr)
	pt, fr := cArray(f.textures)
	frees = append(frees, fr)
	pi, fr := cArray(cimgs)
	frees = append(frees, fr)
	pq, fr := cArray(f.quads)
	frees = append(frees, fr)
	psi, fr := cArray(f.sphereIDs)
	frees = append(frees, fr)
	pqi, fr := cArray(f.quadIDs)
	frees = append(frees, fr)
	ppl, fr := cArray(f.perlins)
	frees = append(frees, fr)

	desc := C.rt_scene_desc{
		abi_version: C.RT_B200_ABI_VERSION,
		spheres:     (*C.rt_sphere)(ps), n_spheres: C.uint64_t(len(f.spheres)),
		materials: (*C.rt_material)(pm), n_materials: C.uint32_t(len(f.materials)),
		textures: (*C.rt_texture)(pt), n_textures: C.uint32_t(len(f.textures)),
		images: (*C.rt_image)(pi), n_images: C.uint32_t(len(cimgs)),
		quads: (*C.rt_quad)(pq), n_quads: C.uint64_t(len(f.quads)),
		sphere_ids: (*C.uint32_t)(psi), quad_ids: (*C.uint32_t)(pqi),
		perlins: (*C.rt_perlin)(ppl), n_perlins: C.uint32_t(len(f.perlins)),
	}
	v3 := func(v Vec3) [3]C.float { return [3]C.float{C.float(v.X), C.float(v.Y), C.float(v.Z)} }
	w, h := int(c.imageWidth), int(c.imageHeight) // camera.go:181-182
	cam := C.rt_camera{
		width: C.int32_t(w), height: C.int32_t(h),
		spp: C.int32_t(c.samplesPerPixel), max_depth: C.int32_t(c.bounceDepth),
		center: v3(c.center), pixel00: v3(c.pixel00), pixel_du: v3(c.pixelDu), pixel_dv: v3(c.pixelDv),
		defocus_u: v3(c.defocusDiskU), defocus_v: v3(c.defocusDiskV),
		defocus_angle: C.float(c.defocusAngleRadians), background: v3(c.background.GetColor()),
	}
	ropts := C.rt_render_opts{seed: C.uint64_t(o.Seed), device: C.int32_t(o.Device)}
	rgb := make([]byte, 3*w*h)
	if len(o.Devices) > 1 { // the whole box from one call
		devs := make([]C.int32_t, len(o.Devices))
		for i, d := range o.Devices {
			devs[i] = C.int32_t(d)
		}
		if o.TileSplit {
			ropts.flags = C.RT_FLAG_TILE_SPLIT
		}
		if err := callB200(func() C.int {
			return C.rt_render_multi(&desc, &cam, &ropts, &devs[0], C.int32_t(len(devs)),
				(*C.uint8_t)(unsafe.Pointer(&rgb[0])), nil, nil)
		}); err != nil {
			return err
		}
	} else {
		var scene *C.rt_scene
		if err := callB200(func() C.int { return C.rt_scene_create(&desc, C.int(o.Device), &scene) }); err != nil {
			return err
		}
		defer C.rt_scene_destroy(scene)
		if err := callB200(func() C.int {
			return C.rt_render(scene, &cam, &ropts, (*C.uint8_t)(unsafe.Pointer(&rgb[0])), nil, nil)
		}); err != nil {
			return err
		}
	}

	if o.PNG {
		img := image.NewNRGBA(image.Rect(0, 0, w, h))
		for p := 0; p < w*h; p++ {
			img.Pix[4*p], img.Pix[4*p+1], img.Pix[4*p+2], img.Pix[4*p+3] = rgb[3*p], rgb[3*p+1], rgb[3*p+2], 255
		}
		return png.Encode(writer, img)
	}
	if o.Binary {
		if _, err := io.WriteString(writer, "P6\n"+strconv.Itoa(w)+" "+strconv.Itoa(h)+"\n255\n"); err != nil {
			return err
		}
		_, err := writer.Write(rgb)
		return err
	}

	// from here on: the reference's own output path (camera.go:183-191, 225, 237-252)
	ppm := []string{"P3", strconv.Itoa(w) + " " + strconv.Itoa(h), "255\n"}
	if _, err := io.WriteString(writer, strings.Join(ppm, "\n")); err != nil {
		return err
	}
	chunks := make(chan []string, 2)
	res := c.StartChunkRenderer(writer, chunks)
	go func() {
		defer close(chunks)
		const agg = 5000 // stage.Agg(ctx.Done(), orderedPixelsOut, 5000)
		for b := 0; b < w*h; b += agg {
			e := b + agg
			if e > w*h {
				e = w * h
			}
			lines := make([]string, 0, e-b)
			for p := b; p < e; p++ {
				lines = append(lines, strconv.Itoa(int(rgb[3*p]))+" "+strconv.Itoa(int(rgb[3*p+1]))+" "+strconv.Itoa(int(rgb[3*p+2])))
			}
			chunks <- lines
		}
	}()
	r := <-res
	if r.err != nil {
		return r.err
	}
	return nil
}
