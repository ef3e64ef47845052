// parity_dump_test.go — pins the B200 build's CPU oracle (and through it the CUDA path) to the REAL reference code.
//
// The reference ships no tests and no golden vectors, and the image the B200 library is built in has no Go toolchain,
// so the oracle (oracle/oracle.cpp, a C++ restatement of this package's hot path) could so far only be checked against
// hand-derived known answers.  This test closes that gap wherever Go exists.  It reads seeded inputs
// (tests/golden/pin_inputs.json of the B200 repository), evaluates them with this package's own functions —
// Sphere.Hit, Quad.Hit, Aabb.Hit, World.Hit, BVH.Hit, Lambertian / Metal / Dielectric / DiffuseLight Scatter and Emit,
// every Texture's GetTexture, ToGamma2 / ToRGB / String, NewCamera's derived state, Camera.GetRay, reflect / refract /
// reflectance and Ray.GetColor — and writes the answers to a JSON fixture.  The B200 repository's
// tests/test_pin_from_go.py then requires the oracle AND the CUDA library to reproduce that fixture bit for bit.
//
// Usage (scripts/pin_from_go.sh of the B200 repository does exactly this):
//
//	cp parity_dump_test.go <reference>/internal/
//	cd <reference> && RT_B200_PIN_INPUTS=<b200>/tests/golden/pin_inputs.json \
//	    RT_B200_PIN_OUTPUTS=<b200>/tests/golden/from_go/pin_outputs.json go test ./internal -run TestB200ParityDump -count=1
//
// It must live in package internal: the fields of HitInfo, ScatterInfo, Camera, Quad and Perlin are unexported.
// Every float32 crosses the JSON files as its IEEE-754 bit pattern (a JSON integer), so nothing is rounded in transit.
//
// Random numbers: the reference draws from math/rand.  Wherever a function takes its *rand.Rand from the caller
// (Lambertian, Metal, GetRay, GetColor) the test hands it a rand.Rand over feedSource, whose Int63 values are chosen so
// that Float32() returns exactly the 24-bit uniforms listed in the input file (Float32() = float32(Float64()),
// Float64() = float64(Int63() & (1<<53 - 1)) / (1<<53), so Int63 = k << 29 yields k / 2^24).  Dielectric.Scatter draws
// from the GLOBAL source (materials.go:103); the test seeds it, records the uniform the call will see, and writes it to
// the fixture so the other side can be fed the same value.
//
// STATUS: written, not compiled here (no Go toolchain in the B200 build image).
package internal

import (
	"encoding/json"
	"image"
	"image/color"
	"math"
	"math/rand"
	"os"
	"path/filepath"
	"testing"
)

// ---- input schema (tests/pin_from_go.py: make_inputs) ---------------------------------------------------------------

type pinRayCase struct {
	C    []uint32 `json:"c"`    // sphere centre
	R    uint32   `json:"r"`    // sphere radius
	Q    []uint32 `json:"q"`    // quad corner
	U    []uint32 `json:"u"`    // quad edge u
	V    []uint32 `json:"v"`    // quad edge v
	Min  []uint32 `json:"min"`  // box
	Max  []uint32 `json:"max"`  // box
	O    []uint32 `json:"o"`    // ray origin
	D    []uint32 `json:"d"`    // ray direction
	Tmin uint32   `json:"tmin"` // interval
	Tmax uint32   `json:"tmax"`
}

type pinTexture struct {
	Kind   string   `json:"kind"` // solid | checker | image | noise
	A      []uint32 `json:"a"`
	B      []uint32 `json:"b"`
	Scale  uint32   `json:"scale"`
	Image  int      `json:"image"`
	Perlin int      `json:"perlin"`
}

type pinMaterial struct {
	Kind    string   `json:"kind"` // lambertian | metal | dielectric | light
	Albedo  []uint32 `json:"albedo"`
	Fuzz    uint32   `json:"fuzz"`
	Ior     uint32   `json:"ior"`
	Texture int      `json:"texture"`
}

type pinImage struct {
	W     int      `json:"w"`
	H     int      `json:"h"`
	RGB16 []uint16 `json:"rgb16"` // row-major, 3 values per texel
}

type pinPerlin struct {
	Vec   [][]uint32 `json:"vec"` // 256 x 3
	PermX []int      `json:"perm_x"`
	PermY []int      `json:"perm_y"`
	PermZ []int      `json:"perm_z"`
}

type pinWorld struct {
	Spheres [][]uint32 `json:"spheres"` // cx, cy, cz, r
	Rays    [][]uint32 `json:"rays"`    // ox, oy, oz, dx, dy, dz
	Tmin    uint32     `json:"tmin"`
	Tmax    uint32     `json:"tmax"`
}

type pinScatterCase struct {
	Material int      `json:"material"`
	Sphere   []uint32 `json:"sphere"` // cx, cy, cz, r
	O        []uint32 `json:"o"`
	D        []uint32 `json:"d"`
	Feed     []uint32 `json:"feed"` // 24-bit uniforms for the ray's rand.Rand, in call order
	Seed     int64    `json:"seed"` // global math/rand seed (Dielectric)
}

type pinTextureCase struct {
	Texture int      `json:"texture"`
	U       uint32   `json:"u"`
	V       uint32   `json:"v"`
	P       []uint32 `json:"p"`
}

type pinCameraCase struct {
	Aspect     uint32   `json:"aspect"`
	Width      int      `json:"width"`
	FovDeg     uint32   `json:"fov_deg"`
	DefocusDeg uint32   `json:"defocus_deg"`
	FocusDist  uint32   `json:"focus_dist"`
	LookFrom   []uint32 `json:"look_from"`
	LookAt     []uint32 `json:"look_at"`
	Background []uint32 `json:"background"`
	Spp        int      `json:"spp"`
	Depth      int      `json:"depth"`
}

type pinGetRayCase struct {
	Camera int      `json:"camera"`
	I      int      `json:"i"`
	J      int      `json:"j"`
	Feed   []uint32 `json:"feed"`
}

type pinVecCase struct {
	A   []uint32 `json:"a"`
	B   []uint32 `json:"b"`
	Eta uint32   `json:"eta"`
	Cos uint32   `json:"cos"`
}

type pinColorScene struct {
	Spheres    [][]uint32 `json:"spheres"`          // cx, cy, cz, r
	SphereMats []int      `json:"sphere_materials"` // material index per sphere
	Quads      [][]uint32 `json:"quads"`            // q(3), u(3), v(3)
	QuadMats   []int      `json:"quad_materials"`
	Background []uint32   `json:"background"`
	Depth      int        `json:"depth"`
	Feed       []uint32   `json:"feed"`
	Rays       [][]uint32 `json:"rays"`
}

type pinInputs struct {
	Version   int              `json:"version"`
	Textures  []pinTexture     `json:"textures"`
	Materials []pinMaterial    `json:"materials"`
	Images    []pinImage       `json:"images"`
	Perlins   []pinPerlin      `json:"perlins"`
	SphereHit []pinRayCase     `json:"sphere_hit"`
	QuadHit   []pinRayCase     `json:"quad_hit"`
	AabbHit   []pinRayCase     `json:"aabb_hit"`
	World     pinWorld         `json:"world"`
	Scatter   []pinScatterCase `json:"scatter"`
	Texture   []pinTextureCase `json:"texture"`
	Resolve   [][]uint32       `json:"resolve"`
	Camera    []pinCameraCase  `json:"camera"`
	GetRay    []pinGetRayCase  `json:"get_ray"`
	Reflect   []pinVecCase     `json:"reflect"`
	Refract   []pinVecCase     `json:"refract"`
	Schlick   []pinVecCase     `json:"reflectance"`
	GetColor  pinColorScene    `json:"get_color"`
}

// ---- output schema ---------------------------------------------------------------------------------------------------

type pinHitOut struct {
	Hit    bool     `json:"hit"`
	T      uint32   `json:"t"`
	Point  []uint32 `json:"point"`
	Normal []uint32 `json:"normal"`
	Front  bool     `json:"front"`
	U      uint32   `json:"u"`
	V      uint32   `json:"v"`
}

type pinIDsOut struct {
	IDs []int    `json:"ids"`
	T   []uint32 `json:"t"`
}

type pinScatterOut struct {
	Hit         bool     `json:"hit"`
	Scattered   bool     `json:"scattered"`
	Origin      []uint32 `json:"origin"`
	Dir         []uint32 `json:"dir"`
	Attenuation []uint32 `json:"attenuation"`
	Emitted     []uint32 `json:"emitted"`
	Uniform     uint32   `json:"uniform"` // what the global rand.Float32() returns first after Seed(seed)
}

type pinResolveOut struct {
	RGB  []int  `json:"rgb"`
	Text string `json:"text"`
}

type pinCameraOut struct {
	Center       []uint32 `json:"center"`
	Pixel00      []uint32 `json:"pixel00"`
	Du           []uint32 `json:"du"`
	Dv           []uint32 `json:"dv"`
	DiskU        []uint32 `json:"disk_u"`
	DiskV        []uint32 `json:"disk_v"`
	Height       int      `json:"height"`
	FovRadians   uint32   `json:"fov_radians"`
	DefocusAngle uint32   `json:"defocus_angle"`
}

type pinRayOut struct {
	Origin []uint32 `json:"origin"`
	Dir    []uint32 `json:"dir"`
}

type pinOutputs struct {
	Producer  string          `json:"producer"`
	Version   int             `json:"version"`
	SphereHit []pinHitOut     `json:"sphere_hit"`
	QuadHit   []pinHitOut     `json:"quad_hit"`
	AabbHit   []bool          `json:"aabb_hit"`
	WorldHit  pinIDsOut       `json:"world_hit"`
	BvhHit    pinIDsOut       `json:"bvh_hit"`
	Scatter   []pinScatterOut `json:"scatter"`
	Texture   [][]uint32      `json:"texture"`
	Resolve   []pinResolveOut `json:"resolve"`
	Camera    []pinCameraOut  `json:"camera"`
	GetRay    []pinRayOut     `json:"get_ray"`
	Reflect   [][]uint32      `json:"reflect"`
	Refract   [][]uint32      `json:"refract"`
	Schlick   []uint32        `json:"reflectance"`
	GetColor  [][]uint32      `json:"get_color"`
}

// ---- helpers ----------------------------------------------------------------------------------------------------------

func pinF(b uint32) float32 { return math.Float32frombits(b) }
func pinB(f float32) uint32 { return math.Float32bits(f) }
func pinV(b []uint32) Vec3  { return NewVec3(pinF(b[0]), pinF(b[1]), pinF(b[2])) }
func pinVB(v Vec3) []uint32 { return []uint32{pinB(v.X), pinB(v.Y), pinB(v.Z)} }

// feedSource makes rand.Rand.Float32() return vals[k] / 2^24, one value per call, the last one for ever after.
type feedSource struct {
	vals []uint32
	pos  int
}

func (s *feedSource) Int63() int64 {
	var v uint32
	if len(s.vals) > 0 {
		if s.pos < len(s.vals) {
			v = s.vals[s.pos]
		} else {
			v = s.vals[len(s.vals)-1]
		}
	}
	s.pos++
	return int64(v) << 29
}

func (s *feedSource) Seed(int64) {}

func pinRand(feed []uint32) *rand.Rand { return rand.New(&feedSource{vals: feed}) }

type pinBuilder struct {
	in       *pinInputs
	images   []image.Image
	perlins  []Perlin
	textures []Texture
}

func newPinBuilder(in *pinInputs) *pinBuilder {
	b := &pinBuilder{in: in}
	for _, im := range in.Images {
		// RGBA64.At returns the stored 16-bit channels unchanged (alpha 0xffff), and the zero colour outside
		// Bounds() — the out-of-bounds colour of these fixtures is therefore (0, 0, 0)
		img := image.NewRGBA64(image.Rect(0, 0, im.W, im.H))
		for y := 0; y < im.H; y++ {
			for x := 0; x < im.W; x++ {
				k := 3 * (y*im.W + x)
				img.SetRGBA64(x, y, color.RGBA64{R: im.RGB16[k], G: im.RGB16[k+1], B: im.RGB16[k+2], A: 0xffff})
			}
		}
		b.images = append(b.images, img)
	}
	for _, p := range in.Perlins {
		per := Perlin{randVec3: make([]Vec3, len(p.Vec)), permX: p.PermX, permY: p.PermY, permZ: p.PermZ}
		for i := range p.Vec {
			per.randVec3[i] = pinV(p.Vec[i])
		}
		b.perlins = append(b.perlins, per)
	}
	for _, t := range in.Textures {
		var tex Texture
		switch t.Kind {
		case "solid":
			a := pinV(t.A)
			tex = NewSolidColor(a.X, a.Y, a.Z)
		case "checker":
			c := NewCheckered(pinF(t.Scale), pinV(t.A), pinV(t.B))
			tex = &c
		case "image":
			it := NewImageTexture(b.images[t.Image])
			tex = &it
		case "noise":
			tex = &NoiseTexture{perlin: b.perlins[t.Perlin], scale: pinF(t.Scale)}
		default:
			panic("unknown texture kind " + t.Kind)
		}
		b.textures = append(b.textures, tex)
	}
	return b
}

func (b *pinBuilder) material(i int) Material {
	m := b.in.Materials[i]
	switch m.Kind {
	case "lambertian":
		l := NewLambertian(b.textures[m.Texture])
		return &l
	case "metal":
		mt := NewMetal(pinV(m.Albedo), pinF(m.Fuzz))
		return &mt
	case "dielectric":
		d := NewDielectric(pinF(m.Ior))
		return &d
	case "light":
		return NewDiffuseLight(b.textures[m.Texture])
	}
	panic("unknown material kind " + m.Kind)
}

func pinHit(hi HitInfo, ok bool) pinHitOut {
	if !ok {
		return pinHitOut{Point: []uint32{0, 0, 0}, Normal: []uint32{0, 0, 0}}
	}
	return pinHitOut{Hit: true, T: pinB(hi.t), Point: pinVB(hi.point), Normal: pinVB(hi.normal), Front: hi.frontFace,
		U: pinB(hi.u), V: pinB(hi.v)}
}

// ---- the dump --------------------------------------------------------------------------------------------------------

func TestB200ParityDump(t *testing.T) {
	inPath := os.Getenv("RT_B200_PIN_INPUTS")
	if inPath == "" {
		inPath = filepath.Join("testdata", "pin_inputs.json")
	}
	outPath := os.Getenv("RT_B200_PIN_OUTPUTS")
	if outPath == "" {
		outPath = filepath.Join("testdata", "pin_outputs.json")
	}
	raw, err := os.ReadFile(inPath)
	if err != nil {
		t.Skipf("no pin inputs (%v): set RT_B200_PIN_INPUTS to tests/golden/pin_inputs.json of the B200 repository", err)
	}
	var in pinInputs
	if err := json.Unmarshal(raw, &in); err != nil {
		t.Fatalf("decoding %s: %v", inPath, err)
	}
	b := newPinBuilder(&in)
	out := pinOutputs{Producer: "go", Version: in.Version}
	dummy := &Lambertian{albedo: NewSolidColor(0.5, 0.5, 0.5)}

	// Sphere.Hit, hittables.go:96-132
	for _, c := range in.SphereHit {
		s := NewSphere(pinV(c.C), pinF(c.R), dummy)
		r := NewRay(pinV(c.O), pinV(c.D), pinRand(nil))
		hi, ok := s.Hit(r, Interval{min: pinF(c.Tmin), max: pinF(c.Tmax)})
		out.SphereHit = append(out.SphereHit, pinHit(hi, ok))
	}
	// Quad.Hit, hittables.go:167-194
	for _, c := range in.QuadHit {
		q := NewQuad(pinV(c.Q), pinV(c.U), pinV(c.V), dummy)
		r := NewRay(pinV(c.O), pinV(c.D), pinRand(nil))
		hi, ok := q.Hit(r, Interval{min: pinF(c.Tmin), max: pinF(c.Tmax)})
		out.QuadHit = append(out.QuadHit, pinHit(hi, ok))
	}
	// Aabb.Hit, bvh.go:52-102
	for _, c := range in.AabbHit {
		a := NewAabb(pinV(c.Min), pinV(c.Max))
		r := NewRay(pinV(c.O), pinV(c.D), pinRand(nil))
		out.AabbHit = append(out.AabbHit, a.Hit(r, Interval{min: pinF(c.Tmin), max: pinF(c.Tmax)}))
	}
	// World.Hit (hittables.go:55-72) and BVH.Hit (bvh.go:220-249) over the same spheres; object ID = insertion index,
	// carried in the red channel of a per-sphere solid colour
	{
		w := NewWorld()
		for i, s := range in.World.Spheres {
			m := &Lambertian{albedo: NewSolidColor(float32(i), 0, 0)}
			w.Add(NewSphere(NewVec3(pinF(s[0]), pinF(s[1]), pinF(s[2])), pinF(s[3]), m))
		}
		tree := NewBVHFromWorld(w)
		idOf := func(hi HitInfo, ok bool) (int, uint32) {
			if !ok {
				return -1, 0
			}
			l := hi.material.(*Lambertian)
			return int(l.albedo.GetTexture(0, 0, NewVec3Zero()).GetColor().X), pinB(hi.t)
		}
		iv := Interval{min: pinF(in.World.Tmin), max: pinF(in.World.Tmax)}
		for _, ray := range in.World.Rays {
			r := NewRay(NewVec3(pinF(ray[0]), pinF(ray[1]), pinF(ray[2])), NewVec3(pinF(ray[3]), pinF(ray[4]), pinF(ray[5])), pinRand(nil))
			hi, ok := w.Hit(r, iv)
			id, tb := idOf(hi, ok)
			out.WorldHit.IDs, out.WorldHit.T = append(out.WorldHit.IDs, id), append(out.WorldHit.T, tb)
			hi, ok = tree.Hit(r, iv)
			id, tb = idOf(hi, ok)
			out.BvhHit.IDs, out.BvhHit.T = append(out.BvhHit.IDs, id), append(out.BvhHit.T, tb)
		}
	}
	// Material.Scatter / Emit, materials.go:33-42, 60-75, 91-113, 301-313, at the hit of one sphere
	for _, c := range in.Scatter {
		mat := b.material(c.Material)
		s := NewSphere(NewVec3(pinF(c.Sphere[0]), pinF(c.Sphere[1]), pinF(c.Sphere[2])), pinF(c.Sphere[3]), mat)
		r := NewRay(pinV(c.O), pinV(c.D), pinRand(c.Feed))
		hi, ok := s.Hit(r, Interval{min: 0.001, max: float32(math.Inf(1))})
		o := pinScatterOut{Hit: ok, Origin: []uint32{0, 0, 0}, Dir: []uint32{0, 0, 0}, Attenuation: []uint32{0, 0, 0},
			Emitted: []uint32{0, 0, 0}}
		if ok {
			rand.Seed(c.Seed)
			o.Uniform = pinB(rand.Float32())
			rand.Seed(c.Seed)
			si, scattered := hi.material.Scatter(r, hi)
			o.Scattered = scattered
			o.Emitted = pinVB(hi.material.Emit(hi.u, hi.v, hi.point).GetColor())
			if scattered {
				o.Origin, o.Dir = pinVB(si.ray.origin), pinVB(si.ray.dir)
				o.Attenuation = pinVB(si.attenuation.GetColor())
			}
		}
		out.Scatter = append(out.Scatter, o)
	}
	// Texture.GetTexture, materials.go:127-137, 155-157, 175-193, 285-288
	for _, c := range in.Texture {
		col := b.textures[c.Texture].GetTexture(pinF(c.U), pinF(c.V), pinV(c.P)).GetColor()
		out.Texture = append(out.Texture, pinVB(col))
	}
	// ToGamma2, ToRGB, String: camera.go:212-214, vec3.go:141-166
	for _, c := range in.Resolve {
		v := pinV(c)
		v.ToGamma2()
		v.ToRGB()
		out.Resolve = append(out.Resolve, pinResolveOut{RGB: []int{int(v.X), int(v.Y), int(v.Z)}, Text: v.String()})
	}
	// NewCamera + init, camera.go:104-166
	cams := make([]*Camera, 0, len(in.Camera))
	for _, c := range in.Camera {
		cam := NewCamera(pinF(c.Aspect), c.Width,
			WithFOVDegrees(pinF(c.FovDeg)), WithDefocusAngleDegrees(pinF(c.DefocusDeg)), WithFocusDist(pinF(c.FocusDist)),
			WithLookFrom(pinV(c.LookFrom)), WithLookAt(pinV(c.LookAt)), WithBackgroundColor(pinV(c.Background)),
			WithSamplesPerPixel(c.Spp), WithMaxRayDepth(c.Depth))
		cams = append(cams, cam)
		out.Camera = append(out.Camera, pinCameraOut{Center: pinVB(cam.center), Pixel00: pinVB(cam.pixel00), Du: pinVB(cam.pixelDu),
			Dv: pinVB(cam.pixelDv), DiskU: pinVB(cam.defocusDiskU), DiskV: pinVB(cam.defocusDiskV), Height: int(cam.imageHeight),
			FovRadians: pinB(cam.fovRadians), DefocusAngle: pinB(cam.defocusAngleRadians)})
	}
	// Camera.GetRay, camera.go:265-299
	for _, c := range in.GetRay {
		cw := &CameraWorker{rand: pinRand(c.Feed)}
		r := cams[c.Camera].GetRay(cw, c.I, c.J)
		out.GetRay = append(out.GetRay, pinRayOut{Origin: pinVB(r.origin), Dir: pinVB(r.dir)})
	}
	// reflect, refract, reflectance: vec3.go:212-221, materials.go:115-119
	for _, c := range in.Reflect {
		out.Reflect = append(out.Reflect, pinVB(reflect(pinV(c.A), pinV(c.B))))
	}
	for _, c := range in.Refract {
		out.Refract = append(out.Refract, pinVB(refract(pinV(c.A), pinV(c.B), pinF(c.Eta))))
	}
	for _, c := range in.Schlick {
		out.Schlick = append(out.Schlick, pinB(reflectance(pinF(c.Cos), pinF(c.Eta))))
	}
	// Ray.GetColor, ray.go:32-54, over a World of spheres and quads without dielectrics (their uniform comes from the
	// global source); with a constant feed the draws do not depend on how many were consumed before
	{
		g := in.GetColor
		w := NewWorld()
		for i, s := range g.Spheres {
			w.Add(NewSphere(NewVec3(pinF(s[0]), pinF(s[1]), pinF(s[2])), pinF(s[3]), b.material(g.SphereMats[i])))
		}
		for i, q := range g.Quads {
			w.Add(NewQuad(pinV(q[0:3]), pinV(q[3:6]), pinV(q[6:9]), b.material(g.QuadMats[i])))
		}
		bg := pinV(g.Background)
		for _, ray := range g.Rays {
			r := NewRay(NewVec3(pinF(ray[0]), pinF(ray[1]), pinF(ray[2])), NewVec3(pinF(ray[3]), pinF(ray[4]), pinF(ray[5])), pinRand(g.Feed))
			out.GetColor = append(out.GetColor, pinVB(r.GetColor(w, bg, g.Depth).GetColor()))
		}
	}

	enc, err := json.Marshal(&out)
	if err != nil {
		t.Fatal(err)
	}
	if err := os.MkdirAll(filepath.Dir(outPath), 0o755); err != nil {
		t.Fatal(err)
	}
	if err := os.WriteFile(outPath, enc, 0o644); err != nil {
		t.Fatal(err)
	}
	t.Logf("wrote %s: %d sphere, %d quad, %d box, %d world rays, %d scatter, %d texture, %d camera, %d GetRay, %d GetColor cases",
		outPath, len(out.SphereHit), len(out.QuadHit), len(out.AabbHit), len(out.WorldHit.IDs), len(out.Scatter),
		len(out.Texture), len(out.Camera), len(out.GetRay), len(out.GetColor))
}
