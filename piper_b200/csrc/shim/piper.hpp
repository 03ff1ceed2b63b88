// piper.hpp-compatible host surface over the piper_b200 engine.
//
// Same namespace, type names, fields and function signatures as the reference's public header
// (/root/reference/src/cpp/piper.hpp:20-128) so that its callers (src/cpp/main.cpp, src/cpp/test.cpp)
// compile against this file unchanged, with two substitutions:
//   * ModelSession holds a `pb200_voice*` (the C ABI of include/piper_b200.h) where the reference
//     holds Ort::Session / Ort::Env / Ort::SessionOptions (piper.hpp:78-85);
//   * text -> phonemes stays on the host behind a hook (PiperConfig::phonemizer) because espeak-ng /
//     piper-phonemize are external to the reference tree (CMakeLists.txt:63-71) and absent here; a
//     host that links piper-phonemize installs phonemize_eSpeak there.  Pre-phonemized input
//     (etc/test_sentences/*.jsonl: "phonemes" / "phoneme_ids") enters through phonemesToAudio /
//     phonemeIdsToAudio, which share all code with textToAudio below the phonemizer call.
#ifndef PIPER_B200_SHIM_PIPER_H_
#define PIPER_B200_SHIM_PIPER_H_

#include <cstdint>
#include <functional>
#include <map>
#include <optional>
#include <ostream>
#include <string>
#include <vector>

#include "mini_json.h"

struct pb200_voice;

namespace piper {

// ---- scalar types ------------------------------------------------------------------------------------------------
using SpeakerId = int64_t;   // row of emb_g; the graph's optional `sid` input
using Phoneme = char32_t;    // one UTF-32 codepoint (piper-phonemize's Phoneme)
using PhonemeId = int64_t;   // the id vector is handed to the engine as int64 without conversion
using json = minijson::Value;

// ---- host-side phonemisation hook --------------------------------------------------------------------------------
struct eSpeakConfig {
  std::string voice = "en-us";   // espeak-ng voice name from "espeak": {"voice": ...}
};

// Turns UTF-8 text into one phoneme vector per sentence.  espeak-ng lives on the host; a program that links
// piper-phonemize assigns its phonemize_eSpeak here.
using Phonemizer =
    std::function<void(const std::string& /*text*/, const eSpeakConfig&, std::vector<std::vector<Phoneme>>& /*sentences*/)>;

struct PiperConfig {
  Phonemizer phonemizer;                            // unset: textToAudio throws, phonemesToAudio still works
  std::string eSpeakDataPath;                       // kept for source compatibility with callers that set it
  std::optional<std::string> tashkeelModelPath;     // idem (libtashkeel is not part of this build)
  bool useESpeak = true;
  bool useTashkeel = false;
};

// ---- what the voice JSON describes -------------------------------------------------------------------------------
enum PhonemeType { eSpeakPhonemes, TextPhonemes };

struct PhonemizeConfig {
  std::map<Phoneme, std::vector<PhonemeId>> phonemeIdMap;               // "phoneme_id_map" (required)
  std::optional<std::map<Phoneme, std::vector<Phoneme>>> phonemeMap;    // "phoneme_map" (rarely used)
  PhonemeType phonemeType = eSpeakPhonemes;                             // "phoneme_type": "text" selects codepoints
  eSpeakConfig eSpeak;
  // id layout produced by phonemes_to_ids: BOS, PAD, (ids, PAD)*, EOS
  PhonemeId idPad = 0, idBos = 1, idEos = 2;
  bool interspersePad = true;
};

struct SynthesisConfig {
  // "inference" block: the three entries of the graph's `scales` input, in that order
  float noiseScale = 0.667f;
  float lengthScale = 1.0f;
  float noiseW = 0.8f;
  // audio format of the produced PCM
  int sampleRate = 22050;   // "audio": {"sample_rate": ...}
  int sampleWidth = 2;      // bytes per sample
  int channels = 1;
  std::optional<SpeakerId> speakerId;                              // set for multi-speaker voices
  std::optional<std::map<Phoneme, float>> phonemeSilenceSeconds;   // "phoneme_silence": split phrases, add silence
  float sentenceSilenceSeconds = 0.2f;                             // appended after every sentence
};

struct ModelConfig {
  std::optional<std::map<std::string, SpeakerId>> speakerIdMap;   // "speaker_id_map"
  int numSpeakers = 1;                                            // "num_speakers"
};

// ---- the engine handle (where the reference keeps Ort::Session / Ort::Env / Ort::SessionOptions) -----------------
struct ModelSession {
  pb200_voice* engine = nullptr;
  int device = 0;
  uint64_t noiseSeed = 0x9E3779B97F4A7C15ull;   // advanced per call: the reference's in-graph noise is unseeded
  ModelSession() = default;
  ~ModelSession();
  ModelSession(const ModelSession&) = delete;
  ModelSession& operator=(const ModelSession&) = delete;
};

struct SynthesisResult {
  double inferSeconds = 0;     // wall time of the engine call(s)
  double audioSeconds = 0;     // samples / sampleRate
  double realTimeFactor = 0;   // inferSeconds / audioSeconds
};

struct Voice {
  json configRoot;
  PhonemizeConfig phonemizeConfig;
  SynthesisConfig synthesisConfig;
  ModelConfig modelConfig;
  ModelSession session;
};

// ---- API ---------------------------------------------------------------------------------------------------------
bool isSingleCodepoint(std::string s);
Phoneme getCodepoint(std::string s);
std::string getVersion();

void initialize(PiperConfig& config);
void terminate(PiperConfig& config);

// `useCuda` selects the CUDA device 0 engine; false is refused (this library has no CPU path).
void loadVoice(PiperConfig& config, std::string modelPath, std::string modelConfigPath, Voice& voice,
               std::optional<SpeakerId>& speakerId, bool useCuda);

// Phoneme ids -> int16 audio appended to audioBuffer (the reference's internal `synthesize`, piper.cpp:337-441)
void synthesize(std::vector<PhonemeId>& phonemeIds, SynthesisConfig& synthesisConfig, ModelSession& session,
                std::vector<int16_t>& audioBuffer, SynthesisResult& result);

// piper-phonemize's phonemes_to_ids: BOS, PAD, (ids(p), PAD)*, EOS; unknown phonemes are counted, not fatal
void phonemes_to_ids(const std::vector<Phoneme>& phonemes, const PhonemizeConfig& config,
                     std::vector<PhonemeId>& phonemeIds, std::map<Phoneme, std::size_t>& missingPhonemes);

void textToAudio(PiperConfig& config, Voice& voice, std::string text, std::vector<int16_t>& audioBuffer,
                 SynthesisResult& result, const std::function<void()>& audioCallback);
void textToWavFile(PiperConfig& config, Voice& voice, std::string text, std::ostream& audioFile,
                   SynthesisResult& result);

// Front doors for pre-phonemized input (same sentence / phrase / silence logic as textToAudio).
void phonemesToAudio(PiperConfig& config, Voice& voice, std::vector<std::vector<Phoneme>>& sentences,
                     std::vector<int16_t>& audioBuffer, SynthesisResult& result,
                     const std::function<void()>& audioCallback);
void phonemeIdsToWavFile(Voice& voice, std::vector<PhonemeId>& phonemeIds, std::ostream& audioFile,
                         SynthesisResult& result);
void writeWavHeader(int sampleRate, int sampleWidth, int channels, uint32_t numSamples, std::ostream& audioFile);

}  // namespace piper

#endif  // PIPER_B200_SHIM_PIPER_H_
