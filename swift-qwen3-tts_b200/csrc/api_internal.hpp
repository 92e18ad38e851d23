// Shared by the translation units that implement the C ABI (api.cu, pool.cu): error plumbing, the opaque handle types and the
// host-buffer decode of a list of utterances on ONE model / GPU.
#pragma once
#include <exception>
#include <new>
#include <string>
#include <vector>

#include "../../include/qwen3tts_cuda.h"
#include "engine.hpp"

namespace q3api {

extern thread_local std::string g_last_error;
int fail(int code, const std::string& msg);   // records the thread-local message, returns `code`

template <typename F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const q3::Error& e) {
    return fail(e.code, e.what());
  } catch (const std::bad_alloc&) {
    return fail(Q3TTS_ENOMEM, "host allocation failed");
  } catch (const std::exception& e) {
    return fail(Q3TTS_EINVAL, e.what());
  }
}

// One utterance of a host-buffer decode: `frames` codec frames whose codes are the contiguous block codes[code_off ...] (frames * Q
// ints, either [T,Q] or [Q,T]) and whose PCM goes to pcm[pcm_off ...] (samples); its audio length goes to lengths[orig].
struct HostUtt { int64_t code_off, pcm_off; int frames, orig; };

// Decode `utts` on model `m` from / into HOST buffers (pageable or pinned).  Takes the model's mutex.  Micro-batches are pipelined:
// the device-to-host copy (and, for pageable destinations, the host-side scatter out of pinned staging) of micro-batch k runs
// under the kernels of micro-batch k+1.  layout: Q3TTS_CODES_BTQ ([T,Q] blocks) or Q3TTS_CODES_BQT ([Q,T] blocks, all of T_uniform
// frames).  Throws q3::Error.
void decode_host_list(q3::Model& m, const int32_t* codes, int32_t layout, int32_t T_uniform, std::vector<HostUtt> utts,
                      void* pcm_out, bool i16, int32_t* lengths_out);

}  // namespace q3api

// A model with open streams outlives q3tts_model_free: the handle is only marked (zombie) and the last q3tts_stream_close deletes it,
// so a stream never dereferences a freed model (the Swift wrapper's deinit order is not under the caller's control).
struct q3tts_model { q3::Model* m = nullptr; int open_streams = 0; bool zombie = false; };
