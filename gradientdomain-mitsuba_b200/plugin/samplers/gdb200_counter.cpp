// Mitsuba 0.5 sampler plugin `gdb200_counter`: the counter-based per-pixel sample stream that the gdb200 tracer and
// its CPU oracle use, as a Sampler plugin, so that a stock Mitsuba build renders the reference `gpt` integrator with
// the SAME random numbers (bit-for-bit comparison of the real reference against libgdb200, SURVEY.md §7 "hard parts").
//
//   <sampler type="gdb200_counter"> <integer name="sampleCount" value="64"/> <integer name="seed" value="0"/> </sampler>
//
// Contract (mirrors gdb200::Sampler in csrc/gpt_device.cuh; the test suite's CPU restatement uses the same generator):
//   * generate(pixel) re-keys a splitmix64 stream from (seed, pixel.x, pixel.y) — the reference's renderBlock calls it
//     exactly once per pixel before that pixel's samples (gpt.cpp:1250-1251), so the stream of a pixel does not depend
//     on which worker thread renders it or in which order (the `independent` sampler's does, independent.cpp:42-45);
//   * next1D()/next2D() consume the stream sequentially over all spp samples of the pixel (gpt never calls advance());
//   * sample arrays (request1DArray/request2DArray) are not used by gpt and are filled from the same stream.
//   * `chunk` (default 0) selects one of gdb200's chunked streams (gdb200_gpt_params.streams_per_pixel = C > 1): chunk 0
//     is the pixel's stream, chunk c > 0 an independently re-keyed one (the key of streamInfo() in csrc/gpt_kernels.cuh).
//     The Sampler API gives no per-sample hook that gpt calls, so a C-stream film is C passes of the reference, pass c with
//     <integer name="chunk" value="c"/> and sampleCount = spp/C (+1 for c < spp%C), films summed (the test suite drives the reference that way).
//
// Built inside a Mitsuba tree.  The test suite compiles it against the reference's real Sampler interface and
// feeds the reference's own gpt.cpp with it; plugin/stub/mitsuba_stub.h is a syntax check where the reference tree is absent.
#if defined(GDB200_STUB_HEADERS)
#include "../stub/mitsuba_stub.h"
#else
#include <mitsuba/render/sampler.h>
#endif
#include <stdint.h>

MTS_NAMESPACE_BEGIN

class GDB200CounterSampler : public Sampler {
public:
	GDB200CounterSampler() : Sampler(Properties()), m_seed(0), m_chunk(0), m_key(0), m_n(0) { }

	GDB200CounterSampler(const Properties &props) : Sampler(props), m_key(0), m_n(0) {
		m_sampleCount = props.getSize("sampleCount", 4);
		m_seed = (uint64_t) props.getSize("seed", 0);
		m_chunk = (uint64_t) props.getSize("chunk", 0);
	}

	GDB200CounterSampler(Stream *stream, InstanceManager *manager) : Sampler(stream, manager), m_key(0), m_n(0) {
		m_seed = stream->readULong();
		m_chunk = stream->readULong();
	}

	void serialize(Stream *stream, InstanceManager *manager) const {
		Sampler::serialize(stream, manager);
		stream->writeULong(m_seed);
		stream->writeULong(m_chunk);
	}

	ref<Sampler> clone() {
		ref<GDB200CounterSampler> sampler = new GDB200CounterSampler();
		sampler->m_sampleCount = m_sampleCount;
		sampler->m_seed = m_seed;
		sampler->m_chunk = m_chunk;
		for (size_t i=0; i<m_req1D.size(); ++i)
			sampler->request1DArray(m_req1D[i]);
		for (size_t i=0; i<m_req2D.size(); ++i)
			sampler->request2DArray(m_req2D[i]);
		return sampler.get();
	}

	static inline uint64_t mix(uint64_t z) {            /* splitmix64 finaliser */
		z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
		z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
		return z ^ (z >> 31);
	}

	void generate(const Point2i &pos) {
		m_key = mix(mix(m_seed + 0x9E3779B97F4A7C15ULL)
			^ ((uint64_t) (uint32_t) pos.x | ((uint64_t) (uint32_t) pos.y << 32)));
		if (m_chunk > 0)
			m_key = mix(m_key ^ (m_chunk * 0xD1B54A32D192ED03ULL));
		m_n = 0;
		for (size_t i=0; i<m_req1D.size(); i++)
			for (size_t j=0; j<m_sampleCount * m_req1D[i]; ++j)
				m_sampleArrays1D[i][j] = next1D();
		for (size_t i=0; i<m_req2D.size(); i++)
			for (size_t j=0; j<m_sampleCount * m_req2D[i]; ++j)
				m_sampleArrays2D[i][j] = next2D();
		m_sampleIndex = 0;
		m_dimension1DArray = m_dimension2DArray = 0;
	}

	Float next1D() {
		++m_n;
		/* 53 random bits -> [0,1): identical in a DOUBLE_PRECISION build, which gpt requires (README.txt:115-118) */
		return (Float) ((double) (mix(m_key + m_n * 0x9E3779B97F4A7C15ULL) >> 11) * (1.0 / 9007199254740992.0));
	}

	Point2 next2D() {
		Float value1 = next1D();
		Float value2 = next1D();
		return Point2(value1, value2);
	}

	std::string toString() const {
		std::ostringstream oss;
		oss << "GDB200CounterSampler[" << endl
			<< "  sampleCount = " << m_sampleCount << "," << endl
			<< "  seed = " << m_seed << "," << endl
			<< "  chunk = " << m_chunk << endl
			<< "]";
		return oss.str();
	}

	MTS_DECLARE_CLASS()
private:
	uint64_t m_seed, m_chunk, m_key, m_n;
};

MTS_IMPLEMENT_CLASS_S(GDB200CounterSampler, false, Sampler)
MTS_EXPORT_PLUGIN(GDB200CounterSampler, "gdb200 counter-based per-pixel sampler");
MTS_NAMESPACE_END
